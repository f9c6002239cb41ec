#!/usr/bin/env python3
"""Transcribe the golden vectors of the reference's own unit tests into tests/golden/reference_vectors.json.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/extract_reference_vectors.py
The JSON is committed; tests only read the JSON.  Sources (all under /root/reference/src):
  encoder.rs:341-620   6 encoder vectors      decoder.rs:256-355   5 decoder block vectors
  bitpacker.rs:196-289 10 bit-packer cases    bitreader.rs:194-304 4 bit-reader scripts
  crc.rs:77-106        2 CRC known answers    x3.rs:200-252        Rice tables (data)
"""
import json
import os
import re

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")


def read(name):
    with open(os.path.join(REF, name)) as f:
        return f.read()


def fn_body(src, fn_name):
    i = src.index("fn %s(" % fn_name)
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        if src[k] == "{":
            depth += 1
        elif src[k] == "}":
            depth -= 1
            if depth == 0:
                return src[j:k + 1]
        k += 1


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def eval_elems(txt, env=None):
    env = env or {}
    txt = strip_comments(txt)
    out = []
    for e in txt.split(","):
        e = e.strip()
        if not e:
            continue
        m = re.fullmatch(r"'(.)' as u8", e)
        if m:
            out.append(ord(m.group(1)))
            continue
        m = re.fullmatch(r"b'(.)'", e)
        if m:
            out.append(ord(m.group(1)))
            continue
        if e in env:
            out.append(env[e])
            continue
        out.append(int(eval(e, {"__builtins__": {}}, {})))  # ints, hex, "-3584 + 11"
    return out


def arrays(body, name):
    """All `let name ... = &[...]` / `&mut [...]` / `[...]` array literals bound to `name`, in order."""
    res = []
    for m in re.finditer(r"let\s+(?:mut\s+)?%s\b[^=]*=\s*(?:&mut\s*|&\s*)?\[" % re.escape(name), body):
        k = m.end()
        depth, j = 1, k
        while depth:
            if body[j] == "[":
                depth += 1
            elif body[j] == "]":
                depth -= 1
            j += 1
        res.append(body[k:j - 1])
    return res


def main():
    v = {}
    enc = read("encoder.rs")
    # --- encoder.rs:342 test_encode_frame, :463 test_encode_frame_zeros
    for fn in ("test_encode_frame", "test_encode_frame_zeros"):
        b = fn_body(enc, fn)
        wav = eval_elems(arrays(b, "wav")[0])
        wl = len(wav)
        exp = eval_elems(arrays(b, "expected_x3_output")[0], {"wlh": wl >> 8, "wll": wl & 0xff})
        v[fn] = {"wav": wav, "expected": exp}
    # --- block level: :494 :520 :566 :595
    for fn, lead in (("test_x3_encode_block", 0), ("test_x3_encode_block_ftype3", 1),
                     ("test_x3_encode_block_bpf_eq16", 0), ("test_x3_encode_block_bpf_lt16", 0)):
        b = fn_body(enc, fn)
        v[fn] = {"wav": eval_elems(arrays(b, "wav")[0]),
                 "expected": eval_elems(arrays(b, "expected_x3_output")[0]),
                 "lead_zero_bits": lead}
    assert "write_packed_zeros(1)" in fn_body(enc, "test_x3_encode_block_ftype3")
    # --- decoder.rs:257-355
    dec = read("decoder.rs")
    for fn in ("test_decode_block_ftype_1", "test_decode_block_ftype_2", "test_decode_block_ftype_3",
               "test_decode_block_bpf_eq16", "test_decode_block_bpf_lt16"):
        b = fn_body(dec, fn)
        item = {"x3_inp": eval_elems(arrays(b, "x3_inp")[0]),
                "expected": eval_elems(arrays(b, "expected_wavput")[0])}
        m = re.search(r"let mut last_wav = (-?\d+);", b)
        if m:  # ftype_1: explicit last_wav, whole buffer is the bit stream, skip 6 bits
            item["last_wav"] = int(m.group(1))
            item["first_sample_prefix"] = False
            item["skip_bits"] = int(re.search(r"br\.read_nbits\((\d+)\);", b).group(1))
        else:  # last_wav = BigEndian i16 at x3_inp[0..2], reader over x3_inp[2..]
            item["first_sample_prefix"] = True
            item["skip_bits"] = 0
        item["wav_len"] = int(re.search(r"&mut \[0i16; (\d+)\]", b).group(1))
        v[fn] = item
    # --- bitpacker.rs:197-289: 10 cases of (init array, [(value, nbits)...], expected)
    bp = fn_body(read("bitpacker.rs"), "test_write_packed_bits")
    cases = []
    for m in re.finditer(r"let inp_arr: &mut \[u8\] = &mut \[([^\]]*)\];(.*?)assert_eq!\(&\[([^\]]*)\], inp_arr\);",
                         bp, re.S):
        writes = [(int(a, 0), int(n)) for a, n in re.findall(r"bp\.write_bits\((0x[0-9a-fA-F]+|\d+)\s*,\s*(\d+)\)", m.group(2))]
        cases.append({"init": eval_elems(m.group(1)), "writes": writes, "expected": eval_elems(m.group(3))})
    assert len(cases) == 10, len(cases)
    v["test_write_packed_bits"] = cases
    # --- crc.rs:78-106
    crc = fn_body(read("crc.rs"), "test_crc")
    hdr = eval_elems(arrays(crc, "header")[0])
    pay = eval_elems(arrays(crc, "payload")[0])
    v["test_crc"] = {"header": hdr, "header_crc_0_16": 0xaddb, "payload": pay, "payload_crc": 2073}
    assert "assert_eq!(0xaddb, crc16(&header[0..16]));" in crc and "assert_eq!(2073, crc16(&payload));" in crc
    # --- x3.rs:200-252 tables
    x3 = read("x3.rs")
    inv = eval_elems(re.search(r"const INV_RICE_CODE: &\[i16\] = &\[(.*?)\];", x3, re.S).group(1))
    tabs = []
    for m in re.finditer(r"RiceCode \{\s*nsubs: (\d+),\s*offset: (\d+),\s*code: &\[(.*?)\],\s*num_bits: &\[(.*?)\],"
                         r"\s*inv: INV_RICE_CODE,\s*inv_len: (\d+),", x3, re.S):
        tabs.append({"nsubs": int(m.group(1)), "offset": int(m.group(2)), "code": eval_elems(m.group(3)),
                     "num_bits": eval_elems(m.group(4)), "inv_len": int(m.group(5))})
    assert len(tabs) == 4
    v["rice_tables"] = {"inv": inv, "codes": tabs}
    with open(OUT, "w") as f:
        json.dump(v, f, separators=(",", ":"))
    print("wrote", OUT, {k: (len(x) if isinstance(x, list) else sorted(x)) for k, x in v.items()})


if __name__ == "__main__":
    main()
