"""TEST INFRASTRUCTURE — generates tests/golden/golden_q15.npz from the UNMODIFIED reference.

    python oracle/gen_golden_q15.py       (build container only: needs oracle/_ref)
Every output array comes from dsps_add_s16_ansi / dsps_mulc_s16_ansi of the real reference
(src/dsp/dsps_add_s16_ansi.c, src/dsp/dsps_mulc_s16_ansi.c) through oracle/_ref; inputs are stored next to them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Reference  # noqa: E402

R = Reference()
rng = np.random.default_rng(20261018)
edge = np.array([0, 1, -1, 2, -2, 32767, -32768, 32766, -32767, 16384, -16384, 255, -256], np.int16)
a = np.concatenate([edge, rng.integers(-32768, 32768, 2000 - edge.size, dtype=np.int64).astype(np.int16)])
b = np.concatenate([edge[::-1], rng.integers(-32768, 32768, 2000 - edge.size, dtype=np.int64).astype(np.int16)])
arrays = {"a": a, "b": b}
cases = []
for k, (n, s1, s2, so, shift) in enumerate([(2000, 1, 1, 1, 0), (2000, 1, 1, 1, 1), (1000, 2, 1, 2, 3), (666, 3, 2, 1, 15),
                                            (13, 1, 1, 1, 0), (0, 1, 1, 1, 0), (500, 1, 4, 3, 16)]):
    out, rc = R.add_s16(a, b, n, s1, s2, so, shift)
    assert rc == 0
    arrays[f"add_{k}"] = out
    cases.append(("add", k, n, s1, s2, so, shift))
for k, (n, c, si, so) in enumerate([(2000, 32767, 1, 1), (2000, -32768, 1, 1), (2000, 16384, 1, 1), (1000, 23170, 2, 1),
                                    (500, -12345, 1, 3), (2000, 0, 1, 1), (2000, 1, 1, 1), (7, 11585, 1, 1)]):
    out, rc = R.mulc_s16(a, n, c, si, so)
    assert rc == 0
    arrays[f"mulc_{k}"] = out
    cases.append(("mulc", k, n, c, si, so, 0))
arrays["cases"] = np.array([[0 if c[0] == "add" else 1] + list(c[1:]) for c in cases], np.int64)
path = os.path.join(ROOT, "tests", "golden", "golden_q15.npz")
np.savez_compressed(path, **arrays)
print("wrote", path, os.path.getsize(path), "bytes")
