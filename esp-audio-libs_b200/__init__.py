"""esp-audio-libs_b200 — B200-native ART resampler path (ctypes binding over the C ABI).

The product is `libesp_audio_b200.so` (hand-written sm_100a kernels behind
include/esp_audio_b200.h); this package is only the thin Python host binding the
tests and bench.py use.  There is no CPU compute path: importing works anywhere (the
`-m "not gpu"` tests check the exported symbols), computing needs a B200.
"""
from .capi import (BLACKMAN_HARRIS, INCLUDE_LOWPASS, MODE_EXACT, MODE_FAST,  # noqa: F401
                   OPT_KERNEL_TIMING, OPT_OVERLAP_STAGING, OPT_PLAN_CACHE, SUBSAMPLE_INTERPOLATE,
                   BiquadBatch, DeviceBuffer, EspbError, PinnedBuffer, ResampleBatch, Resampler, biquad_highpass,
                   biquad_lowpass, checksum_u32, declared_symbols, dsps_add_s16, dsps_mulc_s16, device_count, device_info, float_to_quantized,
                   launch_count, lib, library_path, measure_fp32_fma_peak, measure_fp32_fma_peak2, measure_fp32_tile_pattern, plan_filter_bank, plan_passes, plan_policy,
                   plan_schedule, plan_schedule_segments, quantized_to_float, set_device)
from .capi import measure_host_link, measure_host_link_pattern, ResampleGroups, WavDecoder, wav_write_header, _Q15Backend  # noqa: F401,E402
from .sharding import NcclGather, combine_checksums, gather_words, shard_range  # noqa: F401,E402
