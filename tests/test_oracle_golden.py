"""Pins the CPU oracle against every golden vector the reference's own unit tests hold for the
hot path (transcribed by tests/golden/extract_reference_vectors.py).  CPU only."""
import numpy as np
import pytest


def test_crc_kats(oracle, golden):
    # crc.rs:78-106
    g = golden["test_crc"]
    assert oracle.crc16(bytes(g["header"][:16])) == g["header_crc_0_16"] == 0xaddb
    assert oracle.crc16(bytes(g["payload"])) == g["payload_crc"] == 2073
    assert oracle.crc16(b"123456789") == 0x29B1  # CRC-16/CCITT-FALSE check value


def test_rice_tables_match_reference(oracle, golden):
    # x3.rs:200-252: the closed form used by the oracle (and the CUDA kernels) equals the tables
    g = golden["rice_tables"]
    for k, ref in enumerate(g["codes"]):
        t = oracle.rice_table(k)
        assert t["nsubs"] == ref["nsubs"] and t["offset"] == ref["offset"] and t["inv_len"] == ref["inv_len"]
        assert t["code"] == ref["code"]
        assert t["num_bits"] == ref["num_bits"]
    assert [oracle.inv_rice(i) for i in range(60)] == g["inv"]


def test_bitpacker_cases(oracle, golden):
    # bitpacker.rs:197-289: the packer ORs into a zeroed scratch byte and assigns whole bytes, so bytes it
    # never flushes keep the initial array contents.
    for case in golden["test_write_packed_bits"]:
        buf, n = oracle.bitpack(case["writes"], cap=len(case["init"]))
        got = list(case["init"])
        got[:n] = [int(x) for x in buf[:n]]
        assert got == case["expected"], case


def test_bitreader_scripts(oracle):
    # bitreader.rs:195-211
    (_, lead, rem), = oracle.bitread(bytes([0x00, 0x0f, 0xf0, 0x00]), [-1])
    assert (rem, lead) == (32, 0x000ff000)
    (_, lead, rem), = oracle.bitread(bytes([0x00, 0x0f, 0xf0]), [-1])
    assert (rem, lead) == (24, 0x000ff000)
    # bitreader.rs:213-255 test_count_zero_bits
    r = oracle.bitread(bytes([0x00, 0x0f, 0xf0, 0x00]), [0, 0, 7, 1, 0])
    assert r[0] == (12, 0xff000000, 20)
    assert r[1] == (0, 0xff000000, 20)
    assert r[2] == (0x7f, 0x80000000, 13)
    assert r[3] == (0x01, 0x00000000, 12)
    assert r[4] == (12, 0x00000000, 0)
    # bitreader.rs:257-304 test_bitreader_long_array
    data = bytes([0x01, 0x23, 0x45, 0x67, 0x89, 0xab, 0xcd, 0xef, 0x01])
    r = oracle.bitread(data, [-1, 20, 1, 1, 5, 6, 31, 8])
    assert r[0][1:] == (0b00000001001000110100010101100111, 32)
    assert r[1] == (0b00000001001000110100, 0b010101100111 << 20, 12)
    assert r[2][:2] == (0, 0b10101100111000000000000000000000)
    assert r[3][:2] == (1, 0b01011001110000000000000000000000)
    assert r[4][:2] == (0b01011, 0b00111000000000000000000000000000)
    assert r[5][:2] == (0b001111, 0b00010011010101111001101111011110)
    assert r[6][:2] == (0x09abcdef, 0x01000000)
    assert r[7][:2] == (0x01, 0)


def test_encode_frame_vectors(oracle, golden):
    # encoder.rs:342-460 (1000 samples -> 676 bytes) and :463-491 (zeros)
    for name in ("test_encode_frame", "test_encode_frame_zeros"):
        g = golden[name]
        out, stats = oracle.encode_frame(np.array(g["wav"], dtype=np.int16), cap=0x0eff * 2)
        assert list(out) == g["expected"], name
        assert sum(stats) == len(g["wav"]) - 1
    assert len(golden["test_encode_frame"]["expected"]) == 676


def test_encode_block_vectors(oracle, golden):
    # encoder.rs:494-620
    for name in ("test_x3_encode_block", "test_x3_encode_block_ftype3", "test_x3_encode_block_bpf_eq16",
                 "test_x3_encode_block_bpf_lt16"):
        g = golden[name]
        out = oracle.encode_block_test(np.array(g["wav"], dtype=np.int16), lead_zero_bits=g["lead_zero_bits"])
        assert list(out) == g["expected"], name


def test_decode_block_vectors(oracle, golden):
    # decoder.rs:257-355
    for name in ("test_decode_block_ftype_1", "test_decode_block_ftype_2", "test_decode_block_ftype_3",
                 "test_decode_block_bpf_eq16", "test_decode_block_bpf_lt16"):
        g = golden[name]
        inp = bytes(g["x3_inp"])
        if g["first_sample_prefix"]:
            last = int.from_bytes(inp[:2], "big", signed=True)
            data = inp[2:]
        else:
            last, data = g["last_wav"], inp
        wav = oracle.decode_block_test(data, last, g["wav_len"], skip_bits=g["skip_bits"])
        assert list(wav[:len(g["expected"])]) == g["expected"], name


def test_frame_vector_round_trip_and_header(oracle, golden):
    g = golden["test_encode_frame"]
    frame = bytes(g["expected"])
    h = oracle.read_frame_header(frame[:20])
    assert (h.source_id, h.channels, h.samples, h.payload_len) == (1, 1, 1000, 656)
    assert oracle.crc16(frame[20:]) == h.payload_crc
    pcm = oracle.decode_frame(frame[20:], h.samples)
    assert list(pcm) == g["wav"]
    # mode histogram of this vector: 41 rice3 + 7 BFP + 2 rice1 blocks
    _, stats = oracle.encode_frame(np.array(g["wav"], dtype=np.int16))
    assert stats[0] == 0 and stats[1] == 2 * 20 and stats[4] == 7 * 20 and stats[5] == 0
    assert stats[3] == 40 * 20 + 19
    assert oracle.write_frame_header(1000, 1, 656, h.payload_crc) == frame[:20]
