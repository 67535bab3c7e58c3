// Ratio groups: per-stream / time-varying resampling ratios (ASRC).
//
// In the reference every stream owns a context and `ratio` is an argument of every resampleProcess call
// (art_resampler.cpp:57-62, :167, :208), so two streams may run at different, drifting ratios.  The batched
// kernel shares the position schedule and the expanded coefficients among the streams of a context, which
// needs them in lock-step.  Streams that follow the same clock (same ratio trajectory, same chunking) form a
// *group*; a group set holds one batch context per group and runs the groups of a call concurrently on
// internal CUDA streams forked from / joined to the caller's stream.  Every group keeps its own persistent
// device-side state (history, position), so calls may be chunked freely and each group's ratio may change
// from call to call.
#include <cuda_runtime.h>
#include <stdio.h>

#include <new>
#include <vector>

#include "../../include/esp_audio_b200.h"

struct EspbResampleGroups {
  int channels = 0;
  std::vector<EspbResampleBatch *> ctx;
  std::vector<int> first_stream, n_streams;
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> joined;
  cudaEvent_t fork = nullptr;
};

extern "C" {

void espb_resampleGroupsFree(EspbResampleGroups *g) {
  if (!g)
    return;
  for (EspbResampleBatch *c : g->ctx)
    espb_resampleFree(c);
  for (cudaStream_t s : g->streams)
    cudaStreamDestroy(s);
  for (cudaEvent_t e : g->joined)
    cudaEventDestroy(e);
  if (g->fork)
    cudaEventDestroy(g->fork);
  delete g;
}

EspbResampleGroups *espb_resampleGroupsInit(int num_groups, const int *streams_per_group, int numChannels, int numTaps,
                                            int numFilters, float lowpassRatio, int flags) {
  if (num_groups <= 0 || !streams_per_group) {
    fprintf(stderr, "resampleGroupsInit: needs at least one group\n");
    return nullptr;
  }
  EspbResampleGroups *g = new (std::nothrow) EspbResampleGroups();
  if (!g)
    return nullptr;
  g->channels = numChannels;
  int first = 0;
  for (int k = 0; k < num_groups; ++k) {
    EspbResampleBatch *c = espb_resampleInit(streams_per_group[k], numChannels, numTaps, numFilters, lowpassRatio, flags);
    cudaStream_t s = nullptr;
    cudaEvent_t e = nullptr;
    if (!c || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
      if (c)
        espb_resampleFree(c);
      if (s)
        cudaStreamDestroy(s);
      espb_resampleGroupsFree(g);
      return nullptr;
    }
    g->ctx.push_back(c);
    g->streams.push_back(s);
    g->joined.push_back(e);
    g->first_stream.push_back(first);
    g->n_streams.push_back(streams_per_group[k]);
    first += streams_per_group[k];
  }
  if (cudaEventCreateWithFlags(&g->fork, cudaEventDisableTiming) != cudaSuccess) {
    espb_resampleGroupsFree(g);
    return nullptr;
  }
  return g;
}

int espb_resampleGroupsCount(const EspbResampleGroups *g) { return g ? (int) g->ctx.size() : 0; }

EspbResampleBatch *espb_resampleGroupsContext(EspbResampleGroups *g, int group) {
  return (g && group >= 0 && group < (int) g->ctx.size()) ? g->ctx[group] : nullptr;
}

int espb_resampleGroupsFirstStream(const EspbResampleGroups *g, int group) {
  return (g && group >= 0 && group < (int) g->ctx.size()) ? g->first_stream[group] : -1;
}

int espb_resampleGroupsSetMode(EspbResampleGroups *g, int mode) {
  if (!g)
    return ESPB_ERR_ARG;
  for (EspbResampleBatch *c : g->ctx) {
    const int rc = espb_resampleSetMode(c, mode);
    if (rc != ESPB_OK)
      return rc;
  }
  return ESPB_OK;
}

int espb_resampleGroupsProcessInterleaved(EspbResampleGroups *g, const float *in, int64_t in_stream_stride,
                                          const int *numInputFrames, float *out, int64_t out_stream_stride,
                                          const int *numOutputFrames, const float *ratios, EspbResampleResult *results,
                                          void *stream) {
  if (!g || !numInputFrames || !numOutputFrames || !ratios || !results)
    return ESPB_ERR_ARG;
  cudaStream_t caller = reinterpret_cast<cudaStream_t>(stream);
  if (cudaEventRecord(g->fork, caller) != cudaSuccess)
    return ESPB_ERR_CUDA;
  int rc = ESPB_OK;
  for (size_t k = 0; k < g->ctx.size(); ++k) {
    cudaStream_t s = g->streams[k];
    if (cudaStreamWaitEvent(s, g->fork, 0) != cudaSuccess)
      return ESPB_ERR_CUDA;
    const int64_t first = g->first_stream[k];
    results[k] = espb_resampleProcessInterleaved(g->ctx[k], in + first * in_stream_stride, in_stream_stride,
                                                 numInputFrames[k], out + first * out_stream_stride, out_stream_stride,
                                                 numOutputFrames[k], ratios[k], s);
    if (espb_last_status() != ESPB_OK)
      rc = espb_last_status();
    if (cudaEventRecord(g->joined[k], s) != cudaSuccess || cudaStreamWaitEvent(caller, g->joined[k], 0) != cudaSuccess)
      return ESPB_ERR_CUDA;
  }
  return rc;
}

}  // extern "C"
