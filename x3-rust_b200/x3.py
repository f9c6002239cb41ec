"""Mirror of the reference's `x3` module (src/x3.rs): Parameters, Channel, IterChannel, format constants."""
import ctypes as C

from . import _lib, error


class Archive:                      # x3.rs:136-141
    ID = b"X3ARCHIV"
    ID_LEN = 8


class Frame:                        # x3.rs:143-146
    MAX_LENGTH = 0x7fe0


class FrameHeader:                  # x3.rs:147-184
    LENGTH = 20
    KEY = 30771
    KEY_BUF = b"x3"
    P_KEY, P_SOURCE_ID, P_CHANNELS, P_SAMPLES, P_PAYLOAD_SIZE, P_TIME = 0, 2, 3, 4, 6, 8
    P_HEADER_CRC, P_PAYLOAD_CRC = 16, 18

    def __init__(self, source_id, samples, channels, payload_len, payload_crc):
        self.source_id, self.samples, self.channels = source_id, samples, channels
        self.payload_len, self.payload_crc = payload_len, payload_crc

    def __repr__(self):
        return "FrameHeader(source_id=%d, samples=%d, channels=%d, payload_len=%d, payload_crc=0x%04x)" % (
            self.source_id, self.samples, self.channels, self.payload_len, self.payload_crc)


class Parameters:
    """x3::Parameters (x3.rs:81-134)."""
    MAX_BLOCK_LENGTH = 60
    WAV_BIT_SIZE = 16
    DEFAULT_BLOCK_LENGTH = 20
    DEFAULT_RICE_CODES = (0, 1, 3)
    DEFAULT_THRESHOLDS = (3, 8, 20)
    DEFAULT_BLOCKS_PER_FRAME = 500

    def __init__(self, block_len, blocks_per_frame, codes, thresholds, _validate=True):
        """Parameters::new (x3.rs:99-122): raises X3Error(InvalidEncodingThresh) like the reference."""
        self.block_len = int(block_len)
        self.blocks_per_frame = int(blocks_per_frame)
        self.codes = tuple(int(c) for c in codes)
        self.thresholds = tuple(int(t) for t in thresholds)
        if len(self.codes) != 3 or len(self.thresholds) != 3:
            raise error.X3Error(error.INVALID_ARGUMENT, "codes and thresholds have three entries")
        if _validate:
            error.check(_lib.lib().x3_params_validate(C.byref(self.c_struct())))

    @classmethod
    def default(cls):
        """Parameters::default() (x3.rs:124-134)."""
        return cls(cls.DEFAULT_BLOCK_LENGTH, cls.DEFAULT_BLOCKS_PER_FRAME, cls.DEFAULT_RICE_CODES,
                   cls.DEFAULT_THRESHOLDS, _validate=False)

    def c_struct(self):
        p = _lib.x3_params()
        p.block_len, p.blocks_per_frame = self.block_len, self.blocks_per_frame
        p.codes[:] = self.codes
        p.thresholds[:] = self.thresholds
        return p

    @property
    def samples_per_frame(self):
        return self.block_len * self.blocks_per_frame

    def __repr__(self):
        return "Parameters(block_len=%d, blocks_per_frame=%d, codes=%r, thresholds=%r)" % (
            self.block_len, self.blocks_per_frame, self.codes, self.thresholds)


class Channel:
    """x3::Channel (x3.rs:29-45): slice-backed channel -- the natural GPU entry (contiguous PCM)."""

    def __init__(self, id, wav, sample_rate, params):
        self.id, self.wav, self.sample_rate, self.params = id, wav, sample_rate, params


class IterChannel:
    """x3::IterChannel (x3.rs:47-69): iterator-backed channel.  The GPU path needs contiguous samples, so the
    iterator is drained into an int16 array when encode() is called."""

    def __init__(self, id, wav, sample_rate, params):
        self.id, self.wav, self.sample_rate, self.params = id, iter(wav), sample_rate, params


class X3aSpec:                      # x3.rs:70-79
    def __init__(self, sample_rate, params, channels):
        self.sample_rate, self.params, self.channels = sample_rate, params, channels
