"""Few series — the reference's own use: one context, one stereo (or 8-channel) stream (include/resampler.h:64,
art_resampler.cpp:208-243).  Contexts with <= 32 series run the lanes-over-time kernel (resample_fs_kernel.cu);
ESPB_FS=0 sends the same calls through the standard lanes-over-series kernel.  Both must give the reference's bits in
exact mode — checked here at the full BASELINE sizes against the digests gen_golden.py froze from the unmodified
reference (C1: 10 s of 44.1 -> 48 kHz stereo = 479880 frames, C3/C4/C5 units), and against the oracle on seeded
cases (channel counts 1..8, 1..9 streams, all flag combinations, strong down-sampling, chunked calls, planar I/O)."""
import hashlib

import numpy as np
import pytest
from conftest import bits_equal
from oracle_lib import multitone, noise

import esp_audio_libs_b200 as espb

pytestmark = pytest.mark.gpu
f32 = np.float32
TOL = 1e-6  # BASELINE.json north_star: max-abs 1e-6 full scale (fast mode); exact mode is bit-exact


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert espb.device_count() > 0, "GPU tests need a GPU: the product has no CPU path"
    espb.set_device(0)


@pytest.mark.parametrize("fs", ["1", "0"])
@pytest.mark.parametrize("name", ["C1_10s", "C3_unit_10s", "C4_unit_1s", "C5_unit_10s"])
def test_single_stream_full_size_golden_digest(golden, name, fs, monkeypatch):
    """One stream, one call, exact mode: SHA-256 of the output == the digest of the unmodified reference's output
    (tests/golden/golden_v1.json 'large'), plus frame counts and the final (outputOffset, inputIndex)."""
    monkeypatch.setenv("ESPB_FS", fs)
    _, meta = golden
    c = [m for m in meta["large"] if m["name"] == name][0]
    x = noise(c["n_in"], c["channels"], stream=1, amp=0.5)
    assert sha(x) == c["x_sha256"]
    b = espb.ResampleBatch(1, c["channels"], c["taps"], c["filters"], c["lowpass"], c["flags"], mode=espb.MODE_EXACT)
    b.advance(c["taps"] / 2)
    y, used, gen = b.process_interleaved(x.reshape(1, -1), c["cap"], f32(c["ratio"]))
    assert (used, gen) == (c["used"], c["generated"])
    off, idx = b.state()
    assert (float(off), idx) == (c["final_offset"], c["final_index"])
    assert [float(v) for v in y[0][:8]] == c["y_head"]
    assert sha(y[0][: gen * c["channels"]]) == c["y_sha256"], name
    if fs == "1":  # fast mode on the same call: within the tolerance of the exact result (== the reference)
        bf = espb.ResampleBatch(1, c["channels"], c["taps"], c["filters"], c["lowpass"], c["flags"])
        bf.advance(c["taps"] / 2)
        yf, _, gf = bf.process_interleaved(x.reshape(1, -1), c["cap"], f32(c["ratio"]))
        assert gf == gen
        assert float(np.max(np.abs(yf.astype(np.float64) - y))) <= TOL
        bf.free()
    b.free()


def test_single_stream_c1_chunked_equals_one_shot(golden):
    """C1 in real-time sized calls (441 frames in, state carried on the device): same digest as the one-shot call."""
    _, meta = golden
    c = [m for m in meta["large"] if m["name"] == "C1_10s"][0]
    n_in = 44100 * 2  # two seconds in 10 ms calls, compared with the head of a one-shot call over the same frames
    x = noise(c["n_in"], 2, stream=1, amp=0.5)[: n_in * 2]
    ratio = f32(c["ratio"])
    one = espb.ResampleBatch(1, 2, 256, 256, 1.0, 3, mode=espb.MODE_EXACT)
    one.advance(128)
    y1, u1, g1 = one.process_interleaved(x.reshape(1, -1), n_in * 2, ratio)
    b = espb.ResampleBatch(1, 2, 256, 256, 1.0, 3, mode=espb.MODE_EXACT)
    b.advance(128)
    parts, pos = [], 0
    while pos < n_in:
        n = min(441, n_in - pos)
        y, u, g = b.process_interleaved(x[pos * 2:(pos + n) * 2].reshape(1, -1), 600, ratio, n_in=n)
        assert u == n
        parts.append(y[0][: g * 2])
        pos += u
    yc = np.concatenate(parts)
    assert yc.size == g1 * 2 and bits_equal(yc, y1[0][: g1 * 2])
    assert b.state() == one.state()


CASES = [
    # channels, streams, taps, filters, lowpass, flags, ratio, n_in
    (1, 1, 256, 256, 1.0, 3, f32(48000) / f32(44100), 5000),
    (2, 1, 256, 256, 1.0, 3, f32(48000) / f32(44100), 5000),
    (2, 1, 256, 256, float(f32(44100) / f32(48000) * f32(0.96)), 1, f32(44100) / f32(48000), 5000),
    (3, 1, 64, 16, 1.0, 0, f32(1.37), 3000),
    (2, 2, 128, 37, 0.7, 4, f32(0.61), 4000),
    (8, 1, 1024, 256, 0.45, 1, f32(44100) / f32(96000), 9000),
    (2, 4, 256, 256, 1.0, 3, f32(3.0), 1500),
    (5, 3, 32, 1024, 1.0, 2, f32(0.61), 2000),
    (8, 4, 256, 64, 1.0, 1, f32(2.0), 1500),      # 32 series; ratio 2 -> pass-through / single-filter outputs
    (2, 9, 256, 256, 1.0, 3, f32(48000) / f32(44100), 2500),
    (2, 1, 256, 256, 0.2, 1, f32(0.21), 20000),   # strong down-sampling: wide input tiles
    (1, 2, 4, 2, 1.0, 3, f32(2.5), 300),
    (6, 1, 256, 1024, 1.0, 3, f32(1.0001), 3000),
]


@pytest.mark.parametrize("case", CASES)
def test_few_series_vs_oracle(oracle, case):
    ch, ns, taps, filters, lp, flags, ratio, n_in = case
    x = np.stack([noise(n_in, ch, stream=70 + s, amp=0.9) if s % 2 == 0
                  else multitone(n_in, ch, 44100.0, stream=70 + s, amp=0.9) for s in range(ns)])
    cap = int(n_in * float(ratio)) + 40
    ref = []
    for s in range(ns):
        o = oracle.resampler(ch, taps, filters, lp, flags)
        o.advance(taps / 2)
        yo, uo, go = o.process_interleaved(x[s], cap, ratio)
        ref.append(yo)
    ref = np.stack(ref)
    for mode in (espb.MODE_EXACT, espb.MODE_FAST):
        b = espb.ResampleBatch(ns, ch, taps, filters, lp, flags, mode=mode)
        b.advance(taps / 2)
        # three calls of uneven size: history carried on the device between them
        cuts = (0, n_in // 3 + 1, n_in // 3 + 8, n_in)
        parts = []
        for a, e in zip(cuts[:-1], cuts[1:]):
            y, used, gen = b.process_interleaved(x[:, a * ch:e * ch], cap, ratio, n_in=e - a)
            assert used == e - a
            parts.append(y[:, : gen * ch])
        y = np.concatenate(parts, axis=1)
        assert y.shape[1] == go * ch
        if mode == espb.MODE_EXACT:
            assert bits_equal(y, ref[:, : go * ch]), case
        else:
            assert float(np.max(np.abs(y.astype(np.float64) - ref[:, : go * ch]))) <= TOL, case
        b.free()


def test_few_series_planar(golden):
    arrays, meta = golden
    m = meta["planar"]
    b = espb.ResampleBatch(1, m["channels"], m["taps"], m["filters"], 1.0, m["flags"], mode=espb.MODE_EXACT)
    y, used, gen = b.process_planar(arrays["planar_x"][None], 800, f32(m["ratio"]))
    assert (used, gen) == (m["used"], m["generated"]) and bits_equal(y[0], arrays["planar_y"])


@pytest.mark.parametrize("cfg", [
    # streams, channels, src, dst, src_bits, dst_bits, taps
    (1, 2, 44100, 48000, 16, 16, 256),
    (1, 8, 96000, 44100, 24, 24, 1024),
    (3, 1, 16000, 48000, 16, 32, 256),
    (2, 2, 48000, 44100, 32, 24, 256),
])
def test_wrapper_few_streams_bit_exact(oracle, cfg):
    """Resampler wrapper with a single (or a few) streams: PCM in -> pre/post biquad -> PCM out, bytes and clip
    counts identical to the reference pipeline (oracle), over three chunks."""
    ns, ch, sr, dr, sb, db, taps = cfg
    frames = 3000
    rng = np.random.default_rng(ns * 100 + ch)
    nb = (sb + 7) // 8
    raw = rng.integers(0, 256, size=(ns, 3 * frames * ch * nb), dtype=np.uint8)
    cap = int(frames * dr / sr) + 64
    r = espb.Resampler(ns, frames * ch, cap * ch, sr, dr, sb, db, ch, True, True, taps, 256, mode=espb.MODE_EXACT)
    ob = (db + 7) // 8
    if ch > 2:  # the reference wrapper holds filters for two channels only: compose its stages (one call)
        from test_gpu_parity import _composed_oracle
        pol = r.policy()
        seg = np.ascontiguousarray(raw[:, : frames * ch * nb])
        out, res = r.resample(seg, frames, cap, -1.5)
        for s in range(ns):
            q, used, gen, clipped = _composed_oracle(oracle, seg[s], frames, ch, sb, db, taps, 256, pol, cap, -1.5)
            assert (res["frames_used"], res["frames_generated"]) == (used, gen)
            assert bits_equal(out[s][: gen * ch * ob], q[: gen * ch * ob]), (cfg, s)
            assert int(res["clipped_per_stream"][s]) == clipped
        r.free()
        return
    ws = [oracle.wrapper(frames * ch, cap * ch, float(sr), float(dr), sb, db, ch, True, True, taps, 256)
          for _ in range(ns)]
    for k in range(3):
        seg = np.ascontiguousarray(raw[:, k * frames * ch * nb:(k + 1) * frames * ch * nb])
        out, res = r.resample(seg, frames, cap, -1.5)
        for s in range(ns):
            yo, ro = ws[s].resample(seg[s], frames, cap, -1.5)
            assert res["frames_used"] == ro["frames_used"] and res["frames_generated"] == ro["frames_generated"]
            n = ro["frames_generated"] * ch * ob
            assert bits_equal(out[s][:n], yo[:n]), (cfg, k, s)
            assert int(res["clipped_per_stream"][s]) == ro["clipped_samples"]
    r.free()
