"""Random inputs through the host-side assembly / text-parsing entry points (csrc/assembly.cu) built with
AddressSanitizer + UBSan, checked against difflib (the reference's own dependency) on the way.
usage: host_fuzz_assembly.py LIB SEED N"""
import ctypes as C
import difflib
import sys

import numpy as np

lib = C.CDLL(sys.argv[1])
rng = np.random.default_rng(int(sys.argv[2]))
n_iter = int(sys.argv[3])
lib.nd_longest_match.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_int32)]
lib.nd_parse_signal_text.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int16), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
lib.nd_simple_assembly.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int8), C.POINTER(C.c_int32), C.c_int64,
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
lut = np.full(256, -1, np.int8)
for i, ch in enumerate("ACGTM"):
    lut[ord(ch)] = i
    lut[ord(ch.lower())] = i
mism = 0
for it in range(n_iter):
    alpha = "ACGT" if rng.random() < 0.8 else "ACGTMacgtN"
    genome = "".join(rng.choice(list(alpha), int(rng.integers(1, 1500))))
    # overlapping windows with point errors, some empty, some long (autojunk needs >= 200)
    chunks, pos = [], 0
    while pos < len(genome) and len(chunks) < 40:
        ln = int(rng.integers(0, 420))
        seg = list(genome[pos:pos + ln])
        for _ in range(int(rng.integers(0, 4))):
            if seg:
                seg[int(rng.integers(0, len(seg)))] = str(rng.choice(list(alpha)))
        chunks.append("".join(seg))
        pos += max(1, int(ln * rng.random()))
    # longest match vs difflib on consecutive pairs
    for a, b in zip(chunks[:6], chunks[1:7]):
        out = (C.c_int32 * 3)()
        ea, eb = a.encode(), b.encode()
        assert lib.nd_longest_match(ea, len(ea), eb, len(eb), out) == 0
        m = difflib.SequenceMatcher(None, a, b).find_longest_match(0, len(a), 0, len(b))
        want = (m.a, m.b, m.size) if m.size else (len(a), len(b), 0)      # nothing matches: get_matching_blocks()' sentinel
        mism += (out[0], out[1], out[2]) != want
    enc = [c.encode() for c in chunks if c]
    offs = np.zeros(len(enc) + 1, np.int64)
    if enc:
        offs[1:] = np.cumsum([len(x) for x in enc])
    cap = int(offs[-1]) + 2000
    counts = np.zeros((5, cap), np.int32)
    length, err, args = C.c_int64(0), C.c_int32(0), (C.c_int64 * 2)()
    rc = lib.nd_simple_assembly(b"".join(enc) + b"\0", offs.ctypes.data_as(C.POINTER(C.c_int64)), len(enc),
                                lut.ctypes.data_as(C.POINTER(C.c_int8)), counts.ctypes.data_as(C.POINTER(C.c_int32)), cap,
                                C.byref(length), C.byref(err), args)
    assert rc == 0 and 0 <= length.value <= cap
    # signal text: integers, floats, junk, huge numbers, no trailing separator
    toks = []
    for _ in range(int(rng.integers(0, 300))):
        r = rng.random()
        toks.append(str(int(rng.integers(-40000, 40000))) if r < 0.9 else
                    rng.choice(["1.5", "1e3", "-", "+", "abc", "99999999999999999999999", "-0", "+7", "\x00", "7-"]))
    text = rng.choice([" ", "\n", "\t", "  ", "\r\n"]).join(toks).encode()
    outb = np.empty(len(text) // 2 + 1, np.int16)
    cnt, st = C.c_int64(0), C.c_int32(0)
    capn = outb.size if rng.random() < 0.8 else int(rng.integers(0, max(1, outb.size)))
    assert lib.nd_parse_signal_text(text, len(text), outb.ctypes.data_as(C.POINTER(C.c_int16)), capn, C.byref(cnt), C.byref(st)) == 0
    assert 0 <= cnt.value <= capn
print("assembly fuzz: %d iterations, longest-match mismatches vs difflib: %d" % (n_iter, mism))
