"""The oracle port (oracle/art_oracle.c) against the golden fixtures produced by the
UNMODIFIED reference (oracle/gen_golden.py).  CPU only.  Bit-exact everywhere."""
import hashlib

import numpy as np
import pytest
from conftest import bits_equal
from oracle_lib import multitone, noise

f32 = np.float32


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_filter_banks(oracle, golden):
    arrays, meta = golden
    for k, b in enumerate(meta["banks"]):
        ctx = oracle.resampler(1, b["taps"], b["filters"], b["lowpass"], b["flags"])
        bank = ctx.bank()
        assert ctx.flags() == b["eff_flags"]
        assert sha(bank) == b["sha256"], b
        if f"bank{k}" in arrays:
            assert bits_equal(bank, arrays[f"bank{k}"])
        else:
            assert bits_equal(bank[[0, 1, b["filters"] // 2, b["filters"]]], arrays[f"bank{k}_rows"])
    # KAT (3): centre tap is not exactly 1
    b0 = meta["banks"][0]
    assert b0["centre"] == pytest.approx(0.999999881, abs=1e-9) and b0["centre"] != 1.0


def test_small_cases(oracle, golden):
    arrays, meta = golden
    for c in meta["small"]:
        ctx = oracle.resampler(c["channels"], c["taps"], c["filters"], c["lowpass"], c["flags"])
        if c["advance"]:
            ctx.advance(c["advance"])
        y, used, gen = ctx.process_interleaved(arrays[f"small_{c['name']}_x"], c["cap"], f32(c["ratio"]))
        assert (used, gen) == (c["used"], c["generated"]), c["name"]
        assert bits_equal(y, arrays[f"small_{c['name']}_y"]), c["name"]
        off, idx = ctx.state()
        assert (float(off), idx) == (c["final_offset"], c["final_index"])
        assert ctx.position() == c["position"]


def test_chunked_equals_golden_and_oneshot(oracle, golden):
    arrays, meta = golden
    m = meta["chunked"]
    x, plan = arrays["chunked_x"], arrays["chunked_plan"]
    ch = m["channels"]
    ctx = oracle.resampler(ch, m["taps"], m["filters"], 1.0, m["flags"])
    ctx.advance(m["advance"])
    outs, pos = [], 0
    for n_in, n_out, used, gen in plan:
        y, u, g = ctx.process_interleaved(x[pos * ch:(pos + n_in) * ch], int(n_out), f32(m["ratio"]), n_in=int(n_in))
        assert (u, g) == (used, gen)
        outs.append(y)
        pos += u
    y = np.concatenate(outs)
    assert bits_equal(y, arrays["chunked_y"])
    one = oracle.resampler(ch, m["taps"], m["filters"], 1.0, m["flags"])
    one.advance(m["advance"])
    y1, _, g1 = one.process_interleaved(x, y.size // ch, f32(m["ratio"]))
    assert bits_equal(y1, y)  # KAT (7): state carry is exact


def test_planar(oracle, golden):
    arrays, meta = golden
    m = meta["planar"]
    ctx = oracle.resampler(m["channels"], m["taps"], m["filters"], 1.0, m["flags"])
    y, used, gen = ctx.process_planar(arrays["planar_x"], 800, f32(m["ratio"]))
    assert (used, gen) == (m["used"], m["generated"])
    assert bits_equal(y, arrays["planar_y"])


def test_large_cases_counts_and_digests(oracle, golden):
    _, meta = golden
    for c in meta["large"]:
        x = noise(c["n_in"], c["channels"], stream=1, amp=0.5)
        assert sha(x) == c["x_sha256"], "input generator drifted"
        ctx = oracle.resampler(c["channels"], c["taps"], c["filters"], c["lowpass"], c["flags"])
        ctx.advance(c["taps"] / 2)
        assert ctx.expected(c["n_in"], f32(c["ratio"])) == c["generated"]
        y, used, gen = ctx.process_interleaved(x, c["cap"], f32(c["ratio"]))
        assert (used, gen) == (c["used"], c["generated"]), c["name"]
        off, idx = ctx.state()
        assert (float(off), idx) == (c["final_offset"], c["final_index"])
        assert sha(y) == c["y_sha256"], c["name"]
    # KAT (1): 441000 frames at 44.1->48 give 479880, not the ideal 479861
    c1 = [c for c in meta["large"] if c["name"] == "C1_10s"][0]
    assert c1["generated"] == 479880 and c1["final_index"] == 3496
    assert meta["kat"]["inv_ratio_441_48"] == float(f32(1.0) / (f32(48000) / f32(44100)))


def test_quantisers(oracle, golden):
    arrays, meta = golden
    for q in meta["quant"]:
        bits = q["bits"]
        raw = arrays[f"q2f_{bits}_raw"]
        for gain in (0.0, -6.5):
            assert bits_equal(oracle.quantized_to_float(raw, 4096, bits, gain), arrays[f"q2f_{bits}_{gain}"]), bits
        out, clipped = oracle.float_to_quantized(arrays[f"f2q_{bits}_x"], bits)
        assert clipped == q["clipped"]
        assert bits_equal(out, arrays[f"f2q_{bits}_q"]), bits
    out, clipped = oracle.float_to_quantized(
        np.array([0, 1, -1, 0.5 / 32768, -0.5 / 32768, 1.5 / 32768], f32), 16)
    assert list(out.view(np.int16)) == [0, 32767, -32768, 1, 0, 2] and clipped == 1  # KAT (4)
    # KAT (5): byte 2 is sign-extended in the 32-bit branch
    assert float(oracle.quantized_to_float(np.array([0, 0, 0x80, 1], np.uint8), 1, 32)[0]) == 0.00390625
    assert meta["kat"]["q2f32_quirk"] == 0.00390625


def test_biquad(oracle, golden):
    arrays, meta = golden
    x = arrays["biquad_x"]
    for k, b in enumerate(meta["biquad"]):
        c = oracle.biquad_lowpass(b["f"]) if b["kind"] == "lp" else oracle.biquad_highpass(b["f"])
        assert bits_equal(c, arrays[f"biquad_{k}_c"])
        y = x.copy()
        for chn in range(2):
            s0, s1 = oracle.biquad(c, b["gain"]), oracle.biquad(c, b["gain"])
            s0.apply_buffer(y[chn:], 2, n=4000)
            s1.apply_buffer(y[chn:], 2, n=4000)
        assert bits_equal(y, arrays[f"biquad_{k}_y"])
    assert meta["biquad"][0]["hex"][0] == "0x1.028df80000000p-2"  # KAT (6)
    s = oracle.biquad(arrays["biquad_fo_c"], 1.0)
    assert bits_equal(s.apply_buffer(x[:2000].copy(), 1), arrays["biquad_fo_y"])
    # apply_sample == apply_buffer
    c = oracle.biquad_lowpass(0.1)
    a, b = oracle.biquad(c), oracle.biquad(c)
    ys = np.array([a.apply_sample(v) for v in x[:300]], f32)
    assert bits_equal(ys, b.apply_buffer(x[:300].copy(), 1))


def test_wrapper(oracle, golden):
    arrays, meta = golden
    for k, m in enumerate(meta["wrapper"]):
        chn, nb = m["channels"], (m["src_bits"] + 7) // 8
        w = oracle.wrapper(1024 * chn, 4096 * chn, float(m["src_rate"]), float(m["dst_rate"]), m["src_bits"],
                           m["dst_bits"], chn, m["use_filter"], m["interpolate"], m["taps"], m["filters"])
        raw, outs, res = arrays[f"wrap_{k}_raw"], [], []
        for it in range(3):
            y, r = w.resample(raw[it * 1024 * chn * nb:(it + 1) * 1024 * chn * nb], 1024, m["out_free"][it],
                              m["gain_db"])
            outs.append(y)
            res.append([r["frames_used"], r["frames_generated"], r["predicted_frames_used"], r["clipped_samples"]])
        assert np.array_equal(np.array(res, np.int64), arrays[f"wrap_{k}_res"]), m
        assert bits_equal(np.concatenate(outs), arrays[f"wrap_{k}_y"]), m


def test_invalid_init_returns_null(oracle):
    assert oracle.resampler(1, 30, 16, 1.0, 0) is None      # not a multiple of 4
    assert oracle.resampler(1, 0, 16, 1.0, 0) is None
    assert oracle.resampler(1, 1028, 16, 1.0, 0) is None
    assert oracle.resampler(1, 32, 1, 1.0, 0) is None
    assert oracle.resampler(1, 32, 1025, 1.0, 0) is None
