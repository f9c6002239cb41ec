"""Mirror of src/encodefile.rs: wav_to_x3a (archive id + XML header frame + frames)."""
import wave

import numpy as np

from . import _lib, encoder, x3
from .bytewriter import StreamByteWriter


def read_wav_mono16(wav_filename):
    """16-bit mono PCM only, like encodefile.rs:52,55 (which assert)."""
    with wave.open(str(wav_filename), "rb") as w:
        assert w.getsampwidth() == 2, "Can only handle 16 bit data"          # encodefile.rs:52
        assert w.getnchannels() == 1, "Can only handle one channel"          # encodefile.rs:55
        fs = w.getframerate()
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16, copy=False)
    return pcm, fs


def archive_xml(sample_rate, params):
    """The XML string of encodefile.rs:93-117."""
    return ("<X3ARCH PROG=\"x3new.m\" VERSION=\"2.0\" />"
            "<CFG ID=\"0\" FTYPE=\"XML\" />"
            "<CFG ID=\"1\" FTYPE=\"WAV\">"
            "<FS UNIT=\"Hz\">%d</FS>"
            "<SUFFIX>wav</SUFFIX>"
            "<CODEC TYPE=\"X3\" VERS=\"2\">"
            "<BLKLEN>%d</BLKLEN>"
            "<CODES N=\"4\">RICE%d,RICE%d,RICE%d,BFP</CODES>"
            "<FILTER>DIFF</FILTER>"
            "<NBITS>16</NBITS>"
            "<T N=\"3\">%d,%d,%d</T>"
            "</CODEC>"
            "</CFG>") % ((sample_rate, params.block_len) + tuple(params.codes) + tuple(params.thresholds))


def create_archive_header(sample_rate, params):
    """encodefile.rs:82-138: <Archive Id> + frame header(samples 0, id 0) + XML (+ one zero byte if odd)."""
    import ctypes as C
    xml = archive_xml(sample_rate, params).encode("ascii")
    if len(xml) % 2 == 1:
        xml += b"\0"                                   # the pad byte is part of the payload CRC (:123-128)
    buf = (C.c_uint8 * len(xml)).from_buffer_copy(xml)
    crc = _lib.lib().x3_crc16(buf, len(xml))
    return x3.Archive.ID + encoder.write_frame_header(0, 0, len(xml), crc) + xml


def wav_to_x3a(wav_filename, x3a_filename, quiet=False):
    """encodefile::wav_to_x3a (encodefile.rs:48-78).  Always Parameters::default(), like the reference (:57)."""
    pcm, fs = read_wav_mono16(wav_filename)
    params = x3.Parameters.default()
    ch = x3.Channel(0, pcm, fs, params)
    with open(str(x3a_filename), "wb") as f:
        w = StreamByteWriter(f)
        w.write_all(create_archive_header(fs, params))
        encoder.encode([ch], w, quiet=quiet)
