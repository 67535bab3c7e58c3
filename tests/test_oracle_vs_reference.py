"""Oracle port vs the compiled, unmodified reference (oracle/_ref) on seeded random
cases beyond the committed fixtures.  Skipped when oracle/_ref is absent."""
import numpy as np
from conftest import bits_equal
from oracle_lib import noise

f32 = np.float32


def test_resampler_random_configs(oracle, reference):
    rng = np.random.default_rng(99)
    for _ in range(25):
        taps = int(rng.choice([4, 8, 16, 32, 64, 128, 256, 512]))
        filters = int(rng.integers(2, 300))
        flags = int(rng.integers(0, 4))
        lp = float(rng.choice([1.0, 0.9, 0.5, 0.31]))
        ch = int(rng.integers(1, 5))
        ratio = f32(rng.uniform(0.3, 3.5))
        n_in = int(rng.integers(1, 3000))
        x = noise(n_in, ch, stream=int(rng.integers(0, 1000)), amp=0.9)
        a, b = oracle.resampler(ch, taps, filters, lp, flags), reference.resampler(ch, taps, filters, lp, flags)
        adv = float(rng.choice([0.0, taps / 2, 3.25]))
        a.advance(adv)
        b.advance(adv)
        assert bits_equal(a.bank(), b.bank())
        pos = 0
        while pos < n_in:
            ci, co = int(rng.integers(0, 900)), int(rng.integers(0, 1200))
            ci = min(ci, n_in - pos)
            assert a.required(co, ratio) == b.required(co, ratio)
            assert a.expected(ci, ratio) == b.expected(ci, ratio)
            ya, ua, ga = a.process_interleaved(x[pos * ch:(pos + ci) * ch], co, ratio, n_in=ci)
            yb, ub, gb = b.process_interleaved(x[pos * ch:(pos + ci) * ch], co, ratio, n_in=ci)
            assert (ua, ga) == (ub, gb) and bits_equal(ya, yb)
            assert a.state() == b.state() and a.position() == b.position()
            pos += ua
            if ua == 0 and ga == 0 and ci == 0 and co == 0:
                continue


def test_quantisers_random(oracle, reference):
    rng = np.random.default_rng(3)
    for bits in range(1, 33):
        nb = (bits + 7) // 8
        raw = rng.integers(0, 256, size=3000 * nb, dtype=np.uint8)
        g = float(rng.uniform(-12, 6))
        assert bits_equal(oracle.quantized_to_float(raw, 3000, bits, g), reference.quantized_to_float(raw, 3000, bits, g))
        x = (rng.random(3000) * 2.2 - 1.1).astype(f32)
        qa, ca = oracle.float_to_quantized(x, bits)
        qb, cb = reference.float_to_quantized(x, bits)
        assert ca == cb and bits_equal(qa, qb), bits


def test_biquad_random(oracle, reference):
    rng = np.random.default_rng(4)
    for _ in range(20):
        f = float(rng.uniform(0.001, 0.49))
        ca = oracle.biquad_lowpass(f) if rng.random() < 0.5 else oracle.biquad_highpass(f)
        cb = reference.biquad_lowpass(f)
        assert bits_equal(oracle.biquad_lowpass(f), cb)
        assert bits_equal(oracle.biquad_highpass(f), reference.biquad_highpass(f))
        stride = int(rng.integers(1, 5))
        x = noise(2000, stride, stream=7, amp=1.0)
        gain = float(rng.uniform(0.1, 2.0))
        sa, sb = oracle.biquad(ca, gain), reference.biquad(ca, gain)
        assert bits_equal(sa.apply_buffer(x.copy(), stride), sb.apply_buffer(x.copy(), stride))


def test_wrapper_random(oracle, reference):
    rng = np.random.default_rng(5)
    rates = [8000, 16000, 22050, 32000, 44100, 48000, 96000]
    for _ in range(12):
        sr, dr = float(rng.choice(rates)), float(rng.choice(rates))
        sb, db = int(rng.choice([8, 16, 24, 32])), int(rng.choice([8, 16, 24, 32]))
        ch = int(rng.integers(1, 3))
        taps, filters = int(rng.choice([16, 32, 64])), int(rng.choice([16, 64]))
        use_f, interp = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        a = oracle.wrapper(512 * ch, 4096 * ch, sr, dr, sb, db, ch, use_f, interp, taps, filters)
        b = reference.wrapper(512 * ch, 4096 * ch, sr, dr, sb, db, ch, use_f, interp, taps, filters)
        nb = (sb + 7) // 8
        for it in range(4):
            raw = rng.integers(0, 256, size=512 * ch * nb, dtype=np.uint8)
            free = int(rng.integers(1, 4096))
            ya, ra = a.resample(raw, 512, free, -2.0)
            yb, rb = b.resample(raw, 512, free, -2.0)
            assert ra == rb and bits_equal(ya, yb)
