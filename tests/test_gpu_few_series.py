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


@pytest.mark.parametrize("shape", [(40, 1, 2), (9, 4, 3), (5, 2, 1)])   # groups, streams per group, channels
@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_fused_clock_groups_per_stream_ratios(oracle, shape, mode, monkeypatch):
    """SURVEY 8f N1 as written: every stream (or small group of streams) on its own clock inside one set — the ratio is
    per group AND per call, chunk sizes differ per group, state stays on the device (art_resampler.cpp:57-62,167,208).
    The fused form (one staging / schedule / resampler launch for all groups) against one reference-style context per
    stream fed the same calls: bit-exact in exact mode, <= 1e-6 in fast mode; identical to the one-context-per-group
    form; per-group reset."""
    n_groups, spg, ch = shape
    taps, filters, flags = 64, 128, 3
    ns = n_groups * spg
    rng = np.random.default_rng(7)
    base = [f32(rng.choice([48000 / 44100, 44100 / 48000, 2.0, 0.5, 1.0, 1.37])) for _ in range(n_groups)]
    total = 3000
    x = np.stack([noise(total, ch, stream=900 + s, amp=0.7) for s in range(ns)])
    calls = []
    for call in range(7):
        n_in = [int(rng.integers(0, 500)) for _ in range(n_groups)]
        n_out = [int(rng.integers(0, 900)) for _ in range(n_groups)]
        ratios = [f32(base[k] * f32(1.0 + 3e-4 * np.sin(call + k))) for k in range(n_groups)]
        calls.append((n_in, n_out, ratios))

    def run(fused):
        monkeypatch.setenv("ESPB_GROUPS_FUSED", "1" if fused else "0")
        g = espb.ResampleGroups([spg] * n_groups, ch, taps, filters, 1.0, flags,
                                mode=espb.MODE_EXACT if mode == "exact" else espb.MODE_FAST)
        assert g.is_fused() == fused
        for k in range(n_groups):
            g.advance(k, taps / 2)
        pos = [0] * n_groups
        outs, states = [], []
        for n_in, n_out, ratios in calls:
            n_in = [min(n_in[k], total - pos[k]) for k in range(n_groups)]
            row = max(max(n_in), 1) * ch
            xin = np.zeros((ns, row), f32)
            for s in range(ns):
                k = s // spg
                xin[s, : n_in[k] * ch] = x[s, pos[k] * ch:(pos[k] + n_in[k]) * ch]
            y, res = g.process_interleaved(xin, n_in, n_out, ratios)
            outs.append((xin, n_in, y, res))
            for k in range(n_groups):
                pos[k] += res[k][0]
            states.append([g.state(k) for k in range(n_groups)])
        # reset of one group: its next call starts from silence and the initial position again
        g.reset(0)
        g.advance(0, taps / 2)
        n_in = [200] * n_groups
        xin = np.ascontiguousarray(x[:, : 200 * ch])
        y, res = g.process_interleaved(xin, n_in, [400] * n_groups, [base[k] for k in range(n_groups)])
        g.free()
        return outs, states, (y, res)

    fused, st_f, reset_f = run(True)
    plain, st_p, reset_p = run(False)
    orc = [oracle.resampler(ch, taps, filters, 1.0, flags) for _ in range(ns)]
    for o in orc:
        o.advance(taps / 2)
    worst = 0.0
    for c, ((xin, n_in, y, res), (_, _, yp, resp)) in enumerate(zip(fused, plain)):
        assert res == resp
        assert bits_equal(y, yp), c                      # fused == one context per group
        n_out, ratios = calls[c][1], calls[c][2]
        for s in range(ns):
            k = s // spg
            yo, uo, go = orc[s].process_interleaved(xin[s, : n_in[k] * ch], n_out[k], ratios[k], n_in=n_in[k])
            assert res[k] == (uo, go), (c, s)
            got = y[s, : go * ch]
            if mode == "exact":
                assert bits_equal(got, yo), (c, s)
            elif go:
                worst = max(worst, float(np.max(np.abs(got.astype(np.float64) - yo))))
        assert st_f[c] == st_p[c] == [orc[k * spg].state() for k in range(n_groups)]
    assert worst <= TOL
    # after the reset: group 0 equals a fresh context, the others continue
    y, res = reset_f
    assert bits_equal(y, reset_p[0]) and res == reset_p[1]
    fresh = oracle.resampler(ch, taps, filters, 1.0, flags)
    fresh.advance(taps / 2)
    yo, uo, go = fresh.process_interleaved(x[0, : 200 * ch], 400, base[0], n_in=200)
    assert res[0] == (uo, go)
    if mode == "exact":
        assert bits_equal(y[0, : go * ch], yo)


@pytest.mark.parametrize("ns,ch", [(1, 2), (3, 2), (70, 2), (2, 5)])
def test_planar_pointer_tables(oracle, ns, ch):
    """resampleProcess with the reference's own argument form: one separately allocated buffer per (stream, channel)
    plane (include/art_resampler.h:36-37, art_resampler.cpp:167-202).  Bit-exact with the oracle's planar call, over
    two calls (state carried), through the few-series kernel (<= 32 series) and the standard one (140 series)."""
    taps, filters, flags = 64, 32, 3
    ratio = f32(48000) / f32(44100)
    n_in = 1500
    x = [noise(n_in, 1, stream=1200 + q, amp=0.8) for q in range(ns * ch)]
    b = espb.ResampleBatch(ns, ch, taps, filters, 1.0, flags, mode=espb.MODE_EXACT)
    b.advance(taps / 2)
    os_ = [oracle.resampler(ch, taps, filters, 1.0, flags) for _ in range(ns)]
    for o in os_:
        o.advance(taps / 2)
    for a, e in ((0, 700), (700, n_in)):
        ys, used, gen = b.process_planes([p[a:e] for p in x], 1000, ratio)
        for s in range(ns):
            xi = np.stack([x[s * ch + c][a:e] for c in range(ch)], axis=1).reshape(-1)   # interleave for the oracle
            yo, uo, go = os_[s].process_interleaved(xi, 1000, ratio, n_in=e - a)
            assert (used, gen) == (uo, go)
            for c in range(ch):
                assert bits_equal(ys[s * ch + c], yo.reshape(-1, ch)[:go, c]), (s, c)
    b.free()


def test_cpp_shim_runs_on_the_gpu(oracle, tmp_path):
    """include/esp_audio_b200.hpp — the reference's names over the C ABI — compiled into a small C++ program that
    resamples one stereo stream on the device (resampleInit / AdvancePosition / ProcessInterleaved / GetPosition) and
    filters it (biquad_init / apply_buffer); its output file must equal the oracle's bytes."""
    import os
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n_in, cap = 4000, 4500
    x = noise(n_in, 2, stream=77, amp=0.6)
    x.tofile(tmp_path / "in.f32")
    src = tmp_path / "run.cpp"
    src.write_text(r'''
#include <cstdio>
#include <vector>
#include "esp_audio_b200.hpp"
namespace a = esp_audio_libs_b200::art_resampler;
int main(int argc, char **argv) {
  const int n_in = 4000, cap = 4500;
  std::vector<float> x(n_in * 2), y(cap * 2, 0.0f);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(x.data(), 4, x.size(), f) != x.size()) return 2;
  fclose(f);
  espb_set_device(0);
  a::Resample *r = a::resampleInit(1, 2, 256, 256, 1.0f, a::SUBSAMPLE_INTERPOLATE_ | a::BLACKMAN_HARRIS_);
  if (!r) return 3;
  espb_resampleSetMode(r, ESPB_MODE_EXACT);
  a::resampleAdvancePosition(r, 128.0f);
  float *din = (float *) espb_malloc(x.size() * 4), *dout = (float *) espb_malloc(y.size() * 4);
  espb_memcpy_h2d(din, x.data(), x.size() * 4, nullptr);
  espb_memset(dout, 0, y.size() * 4, nullptr);
  a::ResampleResult res = a::resampleProcessInterleaved(r, din, n_in * 2, n_in, dout, cap * 2, cap, 48000.0f / 44100.0f);
  a::BiquadCoefficients c;
  a::biquad_lowpass(&c, 0.2274);
  a::Biquad *b = a::biquad_init(2, 2, &c, 1.0f);
  if (!b || a::biquad_apply_buffer(b, dout, cap * 2, 2, (int) res.output_generated) != 0) return 4;
  espb_memcpy_d2h(y.data(), dout, y.size() * 4, nullptr);
  espb_device_sync();
  f = fopen(argv[2], "wb");
  fwrite(y.data(), 4, res.output_generated * 2, f);
  fclose(f);
  printf("%u %u %.9g\n", res.input_used, res.output_generated, a::resampleGetPosition(r));
  a::biquad_free(b);
  a::resampleFree(r);
  return 0;
}
''')
    exe = tmp_path / "run"
    libdir = os.path.dirname(espb.library_path())
    subprocess.run(["g++", "-std=c++11", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-lesp_audio_b200", f"-Wl,-rpath,{libdir}"], check=True)
    p = subprocess.run([str(exe), str(tmp_path / "in.f32"), str(tmp_path / "out.f32")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    used, gen, position = p.stdout.split()
    o = oracle.resampler(2, 256, 256, 1.0, 3)
    o.advance(128.0)
    yo, uo, go = o.process_interleaved(x, cap, f32(48000) / f32(44100))
    assert (int(used), int(gen)) == (uo, go) and f32(float(position)) == f32(o.position())
    yo = np.ascontiguousarray(yo[: go * 2])
    c = oracle.biquad_lowpass(0.2274)
    for ch in range(2):
        for _ in range(2):
            oracle.biquad(c, 1.0).apply_buffer(yo[ch:], 2, n=go)
    got = np.fromfile(tmp_path / "out.f32", np.float32)
    assert bits_equal(got, yo)
