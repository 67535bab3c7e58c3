// Multi-GPU plumbing of the stream-sharded batch, in C behind the C ABI.
//
// Streams share no state, so a batch is cut into contiguous stream ranges, one per GPU, and nothing crosses GPUs
// on the data path.  The only exchange is a gather of a few 64-bit words per shard (checksum, frames, clipped
// samples, timings) — one ncclAllGather over NVLink.  Two forms:
//   * EspbMulti: ONE process drives all devices (ncclCommInitAll); one context per device, one ncclAllGather per
//     device inside a group call.
//   * EspbDist:  one process per GPU (torchrun-style launch); rank 0 creates an ncclUniqueId, the launcher's
//     rendezvous carries its 128 bytes to the other ranks, every rank calls espb_dist_init.
// NCCL is resolved at run time (dlopen libnccl.so.2): the library still loads — and every single-GPU entry point
// works — on a machine without NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/esp_audio_b200.h"

namespace {

struct NcclApi {
  void *handle = nullptr;
  bool tried = false;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl() {
  static NcclApi api;
  if (!api.tried) {
    api.tried = true;
    // RTLD_NOLOAD first: when the host application (e.g. PyTorch) has already loaded an NCCL, use that copy
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h)
      h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h)
      h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (h) {
      api.handle = h;
#define ESPB_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
      ESPB_SYM(GetVersion, "ncclGetVersion");
      ESPB_SYM(GetUniqueId, "ncclGetUniqueId");
      ESPB_SYM(CommInitRank, "ncclCommInitRank");
      ESPB_SYM(CommInitAll, "ncclCommInitAll");
      ESPB_SYM(CommDestroy, "ncclCommDestroy");
      ESPB_SYM(AllGather, "ncclAllGather");
      ESPB_SYM(GroupStart, "ncclGroupStart");
      ESPB_SYM(GroupEnd, "ncclGroupEnd");
      ESPB_SYM(GetErrorString, "ncclGetErrorString");
#undef ESPB_SYM
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllGather ||
          !api.GroupStart || !api.GroupEnd)
        api.handle = nullptr;
    }
  }
  return api.handle ? &api : nullptr;
}

thread_local char g_multi_error[256] = "";
int multi_fail(int code, const char *what, const char *detail = nullptr) {
  snprintf(g_multi_error, sizeof(g_multi_error), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
  fprintf(stderr, "[espb] %s\n", g_multi_error);
  return code;
}
const char *nccl_err(ncclResult_t r) {
  NcclApi *n = nccl();
  return (n && n->GetErrorString) ? n->GetErrorString(r) : "NCCL error";
}

}  // namespace

struct EspbMulti {
  std::vector<int> devices;
  std::vector<ncclComm_t> comms;
  std::vector<cudaStream_t> streams;
  std::vector<uint64_t *> send, recv;  // device scratch per device
  int max_words = 0;
};

struct EspbDist {
  int rank = 0, world = 1, device = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  uint64_t *send = nullptr, *recv = nullptr;
  int max_words = 0;
};

extern "C" {

int espb_nccl_version(void) {
  NcclApi *n = nccl();
  int v = 0;
  if (!n || !n->GetVersion || n->GetVersion(&v) != ncclSuccess)
    return 0;
  return v;
}

const char *espb_multi_last_error(void) { return g_multi_error; }

void espb_shard_range(int64_t n_streams, int rank, int world, int64_t *first, int64_t *count) {
  // contiguous ranges whose sizes differ by at most one stream
  const int64_t base = world > 0 ? n_streams / world : 0, extra = world > 0 ? n_streams % world : 0;
  if (first)
    *first = rank * base + (rank < extra ? rank : extra);
  if (count)
    *count = base + (rank < extra ? 1 : 0);
}

// ---------------------------------------------------------------- one process, all devices
void espb_multi_free(EspbMulti *m) {
  if (!m)
    return;
  NcclApi *n = nccl();
  int prev = 0;
  cudaGetDevice(&prev);
  for (size_t i = 0; i < m->devices.size(); ++i) {
    cudaSetDevice(m->devices[i]);
    if (i < m->comms.size() && m->comms[i] && n)
      n->CommDestroy(m->comms[i]);
    if (i < m->streams.size() && m->streams[i])
      cudaStreamDestroy(m->streams[i]);
    if (i < m->send.size() && m->send[i])
      cudaFree(m->send[i]);
    if (i < m->recv.size() && m->recv[i])
      cudaFree(m->recv[i]);
  }
  cudaSetDevice(prev);
  delete m;
}

EspbMulti *espb_multi_create(int n_devices, const int *devices) {
  NcclApi *n = nccl();
  if (!n) {
    multi_fail(ESPB_ERR_STATE, "multi_create: libnccl.so.2 not found");
    return nullptr;
  }
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) {
    cudaGetLastError();
    multi_fail(ESPB_ERR_CUDA, "multi_create: no CUDA device");
    return nullptr;
  }
  if (n_devices <= 0)
    n_devices = have;
  EspbMulti *m = new EspbMulti();
  for (int i = 0; i < n_devices; ++i)
    m->devices.push_back(devices ? devices[i] : i);
  for (int d : m->devices)
    if (d < 0 || d >= have) {
      multi_fail(ESPB_ERR_ARG, "multi_create: device index out of range");
      delete m;
      return nullptr;
    }
  m->comms.assign(n_devices, nullptr);
  m->streams.assign(n_devices, nullptr);
  m->send.assign(n_devices, nullptr);
  m->recv.assign(n_devices, nullptr);
  ncclResult_t r = n->CommInitAll(m->comms.data(), n_devices, m->devices.data());
  if (r != ncclSuccess) {
    multi_fail(ESPB_ERR_CUDA, "ncclCommInitAll", nccl_err(r));
    m->comms.assign(n_devices, nullptr);
    espb_multi_free(m);
    return nullptr;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  m->max_words = 64;
  for (int i = 0; i < n_devices; ++i) {
    cudaSetDevice(m->devices[i]);
    if (cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&m->send[i], m->max_words * sizeof(uint64_t)) != cudaSuccess ||
        cudaMalloc(&m->recv[i], (size_t) m->max_words * n_devices * sizeof(uint64_t)) != cudaSuccess) {
      multi_fail(ESPB_ERR_CUDA, "multi_create: device scratch");
      cudaSetDevice(prev);
      espb_multi_free(m);
      return nullptr;
    }
  }
  cudaSetDevice(prev);
  return m;
}

int espb_multi_size(const EspbMulti *m) { return m ? (int) m->devices.size() : 0; }
int espb_multi_device(const EspbMulti *m, int rank) {
  return (m && rank >= 0 && rank < (int) m->devices.size()) ? m->devices[rank] : -1;
}

// Device-side all-gather: device k contributes `words` uint64 at send_dev[k] and receives world*words at recv_dev[k],
// ordered by rank; enqueued on streams[k] (NULL: the handle's own streams).  Asynchronous.
int espb_multi_allgather_u64(EspbMulti *m, const void *const *send_dev, void *const *recv_dev, int words,
                             void *const *streams) {
  NcclApi *n = nccl();
  if (!m || !n || !send_dev || !recv_dev || words <= 0)
    return multi_fail(ESPB_ERR_ARG, "multi_allgather_u64: bad arguments");
  int prev = 0;
  cudaGetDevice(&prev);
  ncclResult_t r = n->GroupStart();
  for (size_t k = 0; r == ncclSuccess && k < m->devices.size(); ++k) {
    cudaSetDevice(m->devices[k]);
    cudaStream_t s = streams ? reinterpret_cast<cudaStream_t>(streams[k]) : m->streams[k];
    r = n->AllGather(send_dev[k], recv_dev[k], (size_t) words, ncclUint64, m->comms[k], s);
  }
  ncclResult_t r2 = n->GroupEnd();
  cudaSetDevice(prev);
  if (r != ncclSuccess || r2 != ncclSuccess)
    return multi_fail(ESPB_ERR_CUDA, "ncclAllGather", nccl_err(r != ncclSuccess ? r : r2));
  return ESPB_OK;
}

// Host convenience: words_per_rank[k*words + i] is word i of shard k (as each device computed it); every device
// gathers all of them over NCCL; gathered[] receives what device 0 holds afterwards.  Synchronous.
int espb_multi_gather_words(EspbMulti *m, const uint64_t *words_per_rank, int words, uint64_t *gathered) {
  if (!m || !words_per_rank || !gathered || words <= 0 || words > m->max_words)
    return multi_fail(ESPB_ERR_ARG, "multi_gather_words: bad arguments (at most 64 words per rank)");
  const int world = (int) m->devices.size();
  int prev = 0;
  cudaGetDevice(&prev);
  for (int k = 0; k < world; ++k) {
    cudaSetDevice(m->devices[k]);
    if (cudaMemcpyAsync(m->send[k], words_per_rank + (size_t) k * words, words * sizeof(uint64_t),
                        cudaMemcpyHostToDevice, m->streams[k]) != cudaSuccess) {
      cudaSetDevice(prev);
      return multi_fail(ESPB_ERR_CUDA, "multi_gather_words: upload");
    }
  }
  cudaSetDevice(prev);
  std::vector<const void *> s(world);
  std::vector<void *> r(world);
  for (int k = 0; k < world; ++k) {
    s[k] = m->send[k];
    r[k] = m->recv[k];
  }
  int rc = espb_multi_allgather_u64(m, s.data(), r.data(), words, nullptr);
  if (rc != ESPB_OK)
    return rc;
  cudaError_t e = cudaSuccess;
  for (int k = 0; k < world && e == cudaSuccess; ++k) {
    cudaSetDevice(m->devices[k]);
    if (k == 0)
      e = cudaMemcpyAsync(gathered, m->recv[0], (size_t) world * words * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                          m->streams[0]);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(m->streams[k]);
  }
  cudaSetDevice(prev);
  if (e != cudaSuccess)
    return multi_fail(ESPB_ERR_CUDA, "multi_gather_words", cudaGetErrorString(e));
  return ESPB_OK;
}

// ---------------------------------------------------------------- one process per GPU
int espb_dist_unique_id(void *id128) {
  NcclApi *n = nccl();
  if (!n || !id128)
    return multi_fail(ESPB_ERR_STATE, "dist_unique_id: libnccl.so.2 not found");
  ncclUniqueId id;
  ncclResult_t r = n->GetUniqueId(&id);
  if (r != ncclSuccess)
    return multi_fail(ESPB_ERR_CUDA, "ncclGetUniqueId", nccl_err(r));
  memcpy(id128, &id, sizeof(id));
  return ESPB_OK;
}

void espb_dist_free(EspbDist *d) {
  if (!d)
    return;
  NcclApi *n = nccl();
  if (d->comm && n)
    n->CommDestroy(d->comm);
  if (d->stream)
    cudaStreamDestroy(d->stream);
  if (d->send)
    cudaFree(d->send);
  if (d->recv)
    cudaFree(d->recv);
  delete d;
}

// The calling process joins the communicator as `rank` of `world` with the CUDA device that is current.
EspbDist *espb_dist_init(const void *id128, int rank, int world) {
  NcclApi *n = nccl();
  if (!n || !id128 || world <= 0 || rank < 0 || rank >= world) {
    multi_fail(ESPB_ERR_ARG, "dist_init: bad arguments or libnccl.so.2 not found");
    return nullptr;
  }
  EspbDist *d = new EspbDist();
  d->rank = rank;
  d->world = world;
  d->max_words = 64;
  cudaGetDevice(&d->device);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclResult_t r = n->CommInitRank(&d->comm, world, id, rank);
  if (r != ncclSuccess) {
    multi_fail(ESPB_ERR_CUDA, "ncclCommInitRank", nccl_err(r));
    d->comm = nullptr;
    espb_dist_free(d);
    return nullptr;
  }
  if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&d->send, d->max_words * sizeof(uint64_t)) != cudaSuccess ||
      cudaMalloc(&d->recv, (size_t) d->max_words * world * sizeof(uint64_t)) != cudaSuccess) {
    multi_fail(ESPB_ERR_CUDA, "dist_init: device scratch");
    espb_dist_free(d);
    return nullptr;
  }
  return d;
}

int espb_dist_rank(const EspbDist *d) { return d ? d->rank : -1; }
int espb_dist_world(const EspbDist *d) { return d ? d->world : 0; }

// All ranks call this with `words` host words each (<= 64); gathered[world*words] is ordered by rank.  The exchange
// is one ncclAllGather of device buffers; doubles as a barrier (it completes only when every rank has entered it).
int espb_dist_allgather_u64(EspbDist *d, const uint64_t *mine, int words, uint64_t *gathered) {
  NcclApi *n = nccl();
  if (!d || !n || !mine || !gathered || words <= 0 || words > d->max_words)
    return multi_fail(ESPB_ERR_ARG, "dist_allgather_u64: bad arguments (at most 64 words per rank)");
  cudaError_t e = cudaMemcpyAsync(d->send, mine, words * sizeof(uint64_t), cudaMemcpyHostToDevice, d->stream);
  if (e != cudaSuccess)
    return multi_fail(ESPB_ERR_CUDA, "dist_allgather_u64: upload", cudaGetErrorString(e));
  ncclResult_t r = n->AllGather(d->send, d->recv, (size_t) words, ncclUint64, d->comm, d->stream);
  if (r != ncclSuccess)
    return multi_fail(ESPB_ERR_CUDA, "ncclAllGather", nccl_err(r));
  e = cudaMemcpyAsync(gathered, d->recv, (size_t) d->world * words * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                      d->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(d->stream);
  if (e != cudaSuccess)
    return multi_fail(ESPB_ERR_CUDA, "dist_allgather_u64", cudaGetErrorString(e));
  return ESPB_OK;
}

int espb_dist_barrier(EspbDist *d) {
  uint64_t one = 1, all[1024];
  if (!d || d->world > 1024)
    return multi_fail(ESPB_ERR_ARG, "dist_barrier: bad handle");
  return espb_dist_allgather_u64(d, &one, 1, all);
}

}  // extern "C"

// ---------------------------------------------------------------- host link probe
// What the host <-> device link gives plain pinned cudaMemcpyAsync — the ceiling of the host-buffer entry points
// (espb_resampleProcessInterleavedHost, espb_resampler_resample_host).  For each of three patterns — H2D only,
// D2H only, both directions at once — every listed device copies `bytes` per direction in slabs (one cudaMemcpyAsync
// per slab, separate streams per direction), all devices concurrently; results are aggregate GB/s per direction
// over all devices, wall clock between device-wide synchronisations, best of `reps`.
// out[6] = {h2d_only, d2h_only, duplex_h2d, duplex_d2h, duplex_sum, seconds_of_the_duplex_run}.
static int measure_host_link_impl(int n_devices, const int *devices, size_t bytes, size_t slab_bytes, int reps,
                                  int only_pattern, double *out) {
  if (!out || bytes == 0)
    return multi_fail(ESPB_ERR_ARG, "measure_host_link: bad arguments");
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) {
    cudaGetLastError();
    return multi_fail(ESPB_ERR_CUDA, "measure_host_link: no CUDA device");
  }
  int prev = 0;
  cudaGetDevice(&prev);
  std::vector<int> devs;
  if (n_devices <= 0) {
    devs.push_back(prev);
  } else {
    for (int i = 0; i < n_devices; ++i)
      devs.push_back(devices ? devices[i] : i);
  }
  for (int d : devs)
    if (d < 0 || d >= have)
      return multi_fail(ESPB_ERR_ARG, "measure_host_link: device index out of range");
  if (slab_bytes == 0 || slab_bytes > bytes)
    slab_bytes = bytes;
  if (reps < 1)
    reps = 1;
  const size_t n = devs.size();
  std::vector<void *> h_in(n, nullptr), h_out(n, nullptr), d_in(n, nullptr), d_out(n, nullptr);
  std::vector<cudaStream_t> s_in(n, nullptr), s_out(n, nullptr);
  cudaError_t e = cudaSuccess;
  for (size_t k = 0; k < n && e == cudaSuccess; ++k) {
    cudaSetDevice(devs[k]);
    e = cudaMallocHost(&h_in[k], bytes);
    if (e == cudaSuccess)
      e = cudaMallocHost(&h_out[k], bytes);
    if (e == cudaSuccess)
      e = cudaMalloc(&d_in[k], bytes);
    if (e == cudaSuccess)
      e = cudaMalloc(&d_out[k], bytes);
    if (e == cudaSuccess)
      e = cudaStreamCreateWithFlags(&s_in[k], cudaStreamNonBlocking);
    if (e == cudaSuccess)
      e = cudaStreamCreateWithFlags(&s_out[k], cudaStreamNonBlocking);
    if (e == cudaSuccess) {
      memset(h_in[k], 1, bytes);
      memset(h_out[k], 0, bytes);
      e = cudaMemset(d_out[k], 2, bytes);
    }
  }
  auto sync_all = [&]() {
    for (size_t k = 0; k < n; ++k) {
      cudaSetDevice(devs[k]);
      cudaError_t r = cudaDeviceSynchronize();
      if (r != cudaSuccess && e == cudaSuccess)
        e = r;
    }
  };
  double best[3] = {0, 0, 0}, secs_duplex = 0.0;
  std::vector<double> seen;  // single-pattern form: the MEDIAN run (other processes drift in and out of step)
  for (int pattern = 0; pattern < 3 && e == cudaSuccess; ++pattern) {
    if (only_pattern >= 0 && pattern != only_pattern)
      continue;
    for (int rep = 0; rep < reps + 1 && e == cudaSuccess; ++rep) {  // first run of each pattern is a warm-up
      sync_all();
      timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      for (size_t off = 0; off < bytes && e == cudaSuccess; off += slab_bytes) {
        const size_t len = off + slab_bytes <= bytes ? slab_bytes : bytes - off;
        for (size_t k = 0; k < n && e == cudaSuccess; ++k) {
          cudaSetDevice(devs[k]);
          if (pattern != 1)
            e = cudaMemcpyAsync((char *) d_in[k] + off, (char *) h_in[k] + off, len, cudaMemcpyHostToDevice, s_in[k]);
          if (pattern != 0 && e == cudaSuccess)
            e = cudaMemcpyAsync((char *) h_out[k] + off, (char *) d_out[k] + off, len, cudaMemcpyDeviceToHost,
                                s_out[k]);
        }
      }
      sync_all();
      clock_gettime(CLOCK_MONOTONIC, &t1);
      const double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
      const double gbs = (double) bytes * n / s / 1e9;  // per direction, all devices
      if (rep > 0 && only_pattern >= 0)
        seen.push_back(gbs);
      if (rep > 0 && gbs > best[pattern]) {
        best[pattern] = gbs;
        if (pattern == 2)
          secs_duplex = s;
      }
    }
  }
  for (size_t k = 0; k < n; ++k) {
    cudaSetDevice(devs[k]);
    if (h_in[k])
      cudaFreeHost(h_in[k]);
    if (h_out[k])
      cudaFreeHost(h_out[k]);
    if (d_in[k])
      cudaFree(d_in[k]);
    if (d_out[k])
      cudaFree(d_out[k]);
    if (s_in[k])
      cudaStreamDestroy(s_in[k]);
    if (s_out[k])
      cudaStreamDestroy(s_out[k]);
  }
  cudaSetDevice(prev);
  if (e != cudaSuccess)
    return multi_fail(ESPB_ERR_CUDA, "measure_host_link", cudaGetErrorString(e));
  if (only_pattern >= 0 && !seen.empty()) {
    for (size_t a = 0; a < seen.size(); ++a)
      for (size_t b = a + 1; b < seen.size(); ++b)
        if (seen[b] < seen[a]) {
          const double t = seen[a];
          seen[a] = seen[b];
          seen[b] = t;
        }
    best[only_pattern] = seen[seen.size() / 2];
  }
  out[0] = best[0];
  out[1] = best[1];
  out[2] = best[2];
  out[3] = best[2];
  out[4] = 2.0 * best[2];
  out[5] = secs_duplex;
  return ESPB_OK;
}

extern "C" int espb_measure_host_link(int n_devices, const int *devices, size_t bytes, size_t slab_bytes, int reps,
                                      double *out) {
  return measure_host_link_impl(n_devices, devices, bytes, slab_bytes, reps, -1, out);
}

// One pattern only (0: H2D alone, 1: D2H alone, 2: both) on the current device, so that several processes — one per
// GPU — can run the SAME pattern at the same time with a barrier between patterns.  *gbs = GB/s per direction.
extern "C" int espb_measure_host_link_pattern(int pattern, size_t bytes, size_t slab_bytes, int reps, double *gbs) {
  if (pattern < 0 || pattern > 2 || !gbs)
    return multi_fail(ESPB_ERR_ARG, "measure_host_link_pattern: bad arguments");
  double out[6] = {0, 0, 0, 0, 0, 0};
  const int rc = measure_host_link_impl(0, nullptr, bytes, slab_bytes, reps, pattern, out);
  *gbs = out[pattern];
  return rc;
}

// ---------------------------------------------------------------- host link probe, prepared form
// Buffers and streams are set up once (page-locking gigabytes takes about a second and differs from process to
// process); the timed runs can then start on a barrier of the caller's, so that one process per GPU measures the link
// while every other process is copying too — not while they are still allocating.
struct EspbLinkProbe {
  int device = 0;
  size_t bytes = 0, slab = 0;
  void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;
};

extern "C" void espb_link_probe_free(EspbLinkProbe *p) {
  if (!p)
    return;
  if (p->h_in)
    cudaFreeHost(p->h_in);
  if (p->h_out)
    cudaFreeHost(p->h_out);
  if (p->d_in)
    cudaFree(p->d_in);
  if (p->d_out)
    cudaFree(p->d_out);
  if (p->s_in)
    cudaStreamDestroy(p->s_in);
  if (p->s_out)
    cudaStreamDestroy(p->s_out);
  delete p;
}

extern "C" EspbLinkProbe *espb_link_probe_create(size_t bytes, size_t slab_bytes) {
  if (bytes == 0)
    return nullptr;
  EspbLinkProbe *p = new EspbLinkProbe();
  p->bytes = bytes;
  p->slab = (slab_bytes == 0 || slab_bytes > bytes) ? bytes : slab_bytes;
  cudaGetDevice(&p->device);
  cudaError_t e = cudaMallocHost(&p->h_in, bytes);
  if (e == cudaSuccess)
    e = cudaMallocHost(&p->h_out, bytes);
  if (e == cudaSuccess)
    e = cudaMalloc(&p->d_in, bytes);
  if (e == cudaSuccess)
    e = cudaMalloc(&p->d_out, bytes);
  if (e == cudaSuccess)
    e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess)
    e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    memset(p->h_in, 1, bytes);
    memset(p->h_out, 0, bytes);
    e = cudaMemset(p->d_out, 2, bytes);
  }
  if (e == cudaSuccess)
    e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    multi_fail(ESPB_ERR_CUDA, "link_probe_create", cudaGetErrorString(e));
    espb_link_probe_free(p);
    return nullptr;
  }
  return p;
}

// One timed run of pattern 0 (H2D alone) / 1 (D2H alone) / 2 (both): `bytes` per direction in slabs, wall clock
// between device synchronisations.  *gbs = GB/s per direction.
extern "C" int espb_link_probe_run(EspbLinkProbe *p, int pattern, double *gbs, double *seconds) {
  if (!p || pattern < 0 || pattern > 2 || !gbs)
    return multi_fail(ESPB_ERR_ARG, "link_probe_run: bad arguments");
  cudaError_t e = cudaDeviceSynchronize();
  timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (size_t off = 0; off < p->bytes && e == cudaSuccess; off += p->slab) {
    const size_t len = off + p->slab <= p->bytes ? p->slab : p->bytes - off;
    if (pattern != 1)
      e = cudaMemcpyAsync((char *) p->d_in + off, (char *) p->h_in + off, len, cudaMemcpyHostToDevice, p->s_in);
    if (pattern != 0 && e == cudaSuccess)
      e = cudaMemcpyAsync((char *) p->h_out + off, (char *) p->d_out + off, len, cudaMemcpyDeviceToHost, p->s_out);
  }
  if (e == cudaSuccess)
    e = cudaDeviceSynchronize();
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (e != cudaSuccess)
    return multi_fail(ESPB_ERR_CUDA, "link_probe_run", cudaGetErrorString(e));
  const double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  *gbs = (double) p->bytes / s / 1e9;
  if (seconds)
    *seconds = s;
  return ESPB_OK;
}
