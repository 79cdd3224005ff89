// peer_comm.cu -- the gradient exchange of the PPO update over NVLink peer memory, fused with the Adam step.
//
// Replaces `MpiAdamOptimizer.compute_gradients` + `apply_gradients`
// (/root/reference/src/rl/windows_workspace/spinup/utils/mpi_tf.py:45-80: Allreduce(SUM) of the flat gradient, divide by the
// number of processes, Adam, parameter Bcast) for ranks that are GPUs of one node.  The buffer is 57 KB (14 k floats + the five
// loss statistics): the exchange is latency, not bandwidth.  One kernel per optimizer step does
//   A  copy this rank's flat gradient (+ the statistics tail, converted from the gradient kernel's double sums) into its slab,
//      the last CTA to finish publishes the step number into every peer's flag word (store over NVLink);
//   B  wait until every peer's number has arrived in the LOCAL flag words;
//   C  read every rank's slab (peer loads), sum in rank order -- the same order on every rank, so all ranks hold bit-identical
//      sums, parameters and early-stop decisions without any broadcast -- write the sum back and apply TF-1 Adam to the
//      optimizer's slice; the CTA that owns the tail runs the KL early-stop test of ppo.py:268-271 (ml4ca_ppo_ctl).
// Slabs are double-buffered by step parity: a rank overwrites half h two steps later, after it has seen every peer's flag of the
// step in between, which a peer only publishes after its previous kernel (the reader of h) has completed.
// The step number lives in the slab (device memory), not in a kernel argument, so the kernel can sit in a CUDA graph that is
// replayed epoch after epoch.  A wait that exceeds kWaitCycles gives up, counts a timeout (ml4ca_peer_comm_status) and lets the
// kernel finish with whatever it has: a dead peer must not hang the GPU.
#include <string.h>

#include <new>

#include "common.h"

namespace ml4ca {

constexpr int kMaxPeers = 16;
constexpr int kHdrBytes = 256;
constexpr long long kWaitCycles = 10000000000ll;   // ~5 s of SM clock

struct PeerHdr {                 // first kHdrBytes of a slab
  uint32_t flags[kMaxPeers];     // flags[r]: last step rank r has published (written by rank r over NVLink)
  uint32_t seq;                  // steps completed by this rank
  uint32_t arrive;               // CTAs of the running kernel that have finished phase A
  uint32_t timeouts;             // waits given up
};

struct PeerArgs {
  uint8_t* slab[kMaxPeers];      // slab[r]: rank r's slab as mapped into this process (slab[rank] = the local allocation)
  int32_t rank, world;
  int64_t cap;                   // floats per half
  float* buf;                    // [n] in: this rank's values, out: the sums
  int64_t n;
  const double* tail_src;        // nullable: buf[tail_off + q] is taken from (float)tail_src[q], q < n_tail
  int64_t tail_off;
  int32_t n_tail;
  // Adam on buf[lo, hi) (lo == hi: exchange only)
  int64_t lo, hi;
  float* params;
  float* m1;
  float* m2;
  float lr, b1, b2, eps, gscale;
  int32_t net, iter, use_ctl;
  float count, kl_limit;
  ml4ca_ppo_ctl* ctl;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_sys(const float* p) {   // peer memory: never from a stale L1 line
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_adam_kernel(const PeerArgs a) {
  const int32_t* c32 = reinterpret_cast<const int32_t*>(a.ctl);
  if (a.use_ctl && c32 != nullptr && c32[0] != 0 && c32[1] < a.iter) return;   // the loop stopped before this iteration (all ranks agree)
  PeerHdr* hdr = reinterpret_cast<PeerHdr*>(a.slab[a.rank]);
  __shared__ uint32_t target_s;
  __shared__ float lr_t_s;
  if (threadIdx.x == 0) target_s = *reinterpret_cast<volatile uint32_t*>(&hdr->seq) + 1u;
  __syncthreads();
  const uint32_t target = target_s;
  const int64_t half = (int64_t)(target & 1u) * a.cap;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // ---- A: publish ---------------------------------------------------------------------------------------------------
  if (i < a.n) {
    float v = a.buf[i];
    if (a.tail_src != nullptr && i >= a.tail_off && i < a.tail_off + a.n_tail) v = (float)a.tail_src[i - a.tail_off];
    reinterpret_cast<float*>(a.slab[a.rank] + kHdrBytes)[half + i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                 // cumulative: covers the CTA's stores ordered before it by the barrier
    const uint32_t prev = atomicAdd(&hdr->arrive, 1u);
    if (prev == gridDim.x - 1) {            // every CTA's part is in the slab and fenced
      hdr->arrive = 0;
      hdr->seq = target;                    // every CTA has read the old value before it arrived
      __threadfence_system();
      // one system fence above orders the slab before the flags; the flag stores themselves are relaxed so that they go out
      // back to back instead of each waiting for the previous one's acknowledgement over NVLink
      for (int p = 0; p < a.world; ++p) st_relaxed_sys(&reinterpret_cast<PeerHdr*>(a.slab[p])->flags[a.rank], target);
    }
  }
  if (threadIdx.x == 32 && a.hi > a.lo) {   // TF-1 Adam's step size, while the flags travel
    const int t = (a.net == 0 ? a.ctl->t_pi : a.ctl->t_v) + a.iter + 1;
    lr_t_s = (float)((double)a.lr * sqrt(1.0 - pow((double)a.b2, (double)t)) / (1.0 - pow((double)a.b1, (double)t)));
  }
  // ---- B: wait for every rank's step number in the local flag words -----------------------------------------------------
  // (after a first timeout the exchange is dead: later steps do not wait again, the caller sees it in ml4ca_peer_comm_status)
  if ((int)threadIdx.x < a.world && *reinterpret_cast<volatile uint32_t*>(&hdr->timeouts) == 0u) {
    long long t0 = 0;
    for (uint32_t polls = 0; (int32_t)(ld_acquire_sys(&hdr->flags[threadIdx.x]) - target) < 0; ++polls) {
      if (polls < 256) continue;            // the common case: the peer is a few microseconds behind
      if (polls == 256) t0 = clock64();
      __nanosleep(200);
      if (clock64() - t0 > kWaitCycles) {
        atomicAdd(&hdr->timeouts, 1u);
        break;
      }
    }
  }
  __syncthreads();
  // ---- C: sum in rank order, write back, Adam ----------------------------------------------------------------------------
  if (i < a.n) {
    // all peer loads are issued before the first sum: one NVLink round trip, not `world` of them
    float v[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      v[p] = p < a.world ? ld_sys(reinterpret_cast<const float*>(a.slab[p] + kHdrBytes) + half + i) : 0.f;
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < a.world) s += v[p];
    a.buf[i] = s;
    if (i >= a.lo && i < a.hi) {
      adam_update(a.params[i], a.m1[i], a.m2[i], __fmul_rn(s, a.gscale), lr_t_s, a.b1, a.b2, a.eps);
    }
  }
  // the thread that summed the approx-KL of the tail runs the bookkeeping of adam_dev_kernel (ppo_update.cu)
  if (a.hi > a.lo && a.ctl != nullptr && a.n_tail >= 5 && i == a.tail_off) {
    float tail[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p0 = 0; p0 < a.world; p0 += 4) {     // four ranks' statistics per round trip, summed in rank order like the rest
      float v[4][5];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
#pragma unroll
        for (int q = 0; q < 5; ++q)
          v[pp][q] = p0 + pp < a.world ? ld_sys(reinterpret_cast<const float*>(a.slab[p0 + pp] + kHdrBytes) + half + a.tail_off + q) : 0.f;
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
#pragma unroll
        for (int q = 0; q < 5; ++q)
          if (p0 + pp < a.world) tail[q] += v[pp][q];
    }
    if (a.iter == 0) {
      if (a.net == 0) {
        for (int q = 0; q < 5; ++q) a.ctl->first[q] = tail[q];
      } else {
        a.ctl->first[5] = tail[1];
      }
    }
    if (a.net == 0 && a.kl_limit > 0.f && tail[2] / a.count > a.kl_limit) {   // ppo.py:269-271; the step above stays applied
      a.ctl->stop_iter = a.iter;
      __threadfence();
      a.ctl->stop = 1;
    }
  }
}

}  // namespace ml4ca

struct ml4ca_peer_comm {
  int32_t rank, world, device;
  int64_t cap;
  uint8_t* slab[ml4ca::kMaxPeers];
  bool opened[ml4ca::kMaxPeers];
  bool connected;
};

using namespace ml4ca;

extern "C" {

int ml4ca_peer_comm_create(int32_t rank, int32_t world, int64_t max_floats, int32_t device, ml4ca_peer_comm** out) {
  ML4CA_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  ML4CA_REQUIRE(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world, "2 <= world <= 16 and 0 <= rank < world");
  ML4CA_REQUIRE(max_floats >= 1, "max_floats must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("ml4ca_peer_comm_create: no CUDA device (this library has no CPU fallback)");
    return ML4CA_ERR_NO_DEVICE;
  }
  ML4CA_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  ML4CA_CUDA(cudaSetDevice(device));
  ml4ca_peer_comm* c = new (std::nothrow) ml4ca_peer_comm();
  ML4CA_REQUIRE(c != nullptr, "out of host memory");
  c->rank = rank, c->world = world, c->device = device, c->connected = false;
  c->cap = (max_floats + 3) / 4 * 4;
  for (int p = 0; p < kMaxPeers; ++p) c->slab[p] = nullptr, c->opened[p] = false;
  const size_t bytes = kHdrBytes + 2 * (size_t)c->cap * sizeof(float);
  void* mem = nullptr;
  int st = check_cuda(cudaMalloc(&mem, bytes), "cudaMalloc(peer slab)");
  if (st != ML4CA_OK) {
    delete c;
    return st;
  }
  st = check_cuda(cudaMemset(mem, 0, bytes), "cudaMemset(peer slab)");
  if (st == ML4CA_OK) st = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  if (st != ML4CA_OK) {
    cudaFree(mem);
    delete c;
    return st;
  }
  c->slab[rank] = static_cast<uint8_t*>(mem);
  *out = c;
  return ML4CA_OK;
}

int ml4ca_peer_comm_export(const ml4ca_peer_comm* c, uint8_t* handle64) {
  ML4CA_REQUIRE(c != nullptr && handle64 != nullptr, "comm and handle are required");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  ML4CA_CUDA(cudaSetDevice(c->device));
  ML4CA_CUDA(cudaIpcGetMemHandle(&h, c->slab[c->rank]));
  memcpy(handle64, &h, 64);
  return ML4CA_OK;
}

int ml4ca_peer_comm_connect(ml4ca_peer_comm* c, const uint8_t* handles) {
  ML4CA_REQUIRE(c != nullptr && handles != nullptr, "comm and handles are required");
  ML4CA_REQUIRE(!c->connected, "already connected");
  ML4CA_CUDA(cudaSetDevice(c->device));
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)p, 64);
    void* mem = nullptr;
    ML4CA_CUDA(cudaIpcOpenMemHandle(&mem, h, cudaIpcMemLazyEnablePeerAccess));
    c->slab[p] = static_cast<uint8_t*>(mem);
    c->opened[p] = true;
  }
  c->connected = true;
  return ML4CA_OK;
}

int ml4ca_peer_comm_slab(const ml4ca_peer_comm* c, void** slab) {
  ML4CA_REQUIRE(c != nullptr && slab != nullptr, "comm and slab are required");
  *slab = c->slab[c->rank];
  return ML4CA_OK;
}

int ml4ca_peer_comm_connect_ptrs(ml4ca_peer_comm* c, void* const* slabs) {
  ML4CA_REQUIRE(c != nullptr && slabs != nullptr, "comm and slabs are required");
  ML4CA_REQUIRE(!c->connected, "already connected");
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    ML4CA_REQUIRE(slabs[p] != nullptr, "a peer slab is NULL");
    c->slab[p] = static_cast<uint8_t*>(slabs[p]);
  }
  c->connected = true;
  return ML4CA_OK;
}

int ml4ca_peer_comm_destroy(ml4ca_peer_comm* c) {
  if (c == nullptr) return ML4CA_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < c->world; ++p)
    if (c->opened[p]) cudaIpcCloseMemHandle(c->slab[p]);
  if (c->slab[c->rank] != nullptr) cudaFree(c->slab[c->rank]);
  delete c;
  return ML4CA_OK;
}

int ml4ca_peer_comm_status(ml4ca_peer_comm* c, int32_t* steps, int32_t* timeouts) {
  ML4CA_REQUIRE(c != nullptr, "comm is NULL");
  ML4CA_CUDA(cudaSetDevice(c->device));
  PeerHdr h;
  ML4CA_CUDA(cudaMemcpy(&h, c->slab[c->rank], sizeof(h), cudaMemcpyDeviceToHost));
  if (steps != nullptr) *steps = (int32_t)h.seq;
  if (timeouts != nullptr) *timeouts = (int32_t)h.timeouts;
  return ML4CA_OK;
}

static int launch_peer(ml4ca_peer_comm* c, PeerArgs& a, cudaStream_t st) {
  ML4CA_REQUIRE(c->connected, "ml4ca_peer_comm_connect has not run");
  ML4CA_REQUIRE(a.n >= 1 && a.n <= c->cap, "buffer longer than the slab");
  for (int p = 0; p < kMaxPeers; ++p) a.slab[p] = c->slab[p];
  a.rank = c->rank, a.world = c->world, a.cap = c->cap;
  const int grid = (int)((a.n + 255) / 256);
  peer_adam_kernel<<<grid, 256, 0, st>>>(a);
  return check_launch("peer_adam_kernel");
}

int ml4ca_peer_allreduce(ml4ca_peer_comm* c, float* buf, int64_t n, const double* tail_src, int64_t tail_off, int32_t n_tail,
                         const ml4ca_ppo_ctl* ctl, int32_t iter, void* stream) {
  ML4CA_REQUIRE(c != nullptr && buf != nullptr, "comm and buf are required");
  ML4CA_REQUIRE(tail_src == nullptr || (tail_off >= 0 && n_tail >= 0 && tail_off + n_tail <= n), "tail outside the buffer");
  PeerArgs a = {};
  a.buf = buf, a.n = n, a.tail_src = tail_src, a.tail_off = tail_off, a.n_tail = tail_src ? n_tail : 0;
  a.ctl = const_cast<ml4ca_ppo_ctl*>(ctl), a.use_ctl = ctl != nullptr, a.iter = iter;
  return launch_peer(c, a, static_cast<cudaStream_t>(stream));
}

int ml4ca_adam_step_peer(ml4ca_peer_comm* c, float* buf, int64_t n, const double* tail_src, int64_t tail_off, int64_t lo, int64_t hi,
                         float* params, float* m1, float* m2, float lr, float beta1, float beta2, float eps, float grad_scale,
                         int32_t net, int32_t iter, float count, float kl_limit, ml4ca_ppo_ctl* ctl, void* stream) {
  ML4CA_REQUIRE(c != nullptr && buf && params && m1 && m2 && ctl, "comm, buf, params, m1, m2 and ctl are required");
  ML4CA_REQUIRE(0 <= lo && lo < hi && hi <= n && iter >= 0 && (net == 0 || net == 1) && count > 0.f, "bad arguments");
  ML4CA_REQUIRE(tail_src != nullptr && tail_off >= hi && tail_off + 5 <= n, "the five statistics live behind the gradient");
  PeerArgs a = {};
  a.buf = buf, a.n = n, a.tail_src = tail_src, a.tail_off = tail_off, a.n_tail = 5;
  a.lo = lo, a.hi = hi, a.params = params, a.m1 = m1, a.m2 = m2;
  a.lr = lr, a.b1 = beta1, a.b2 = beta2, a.eps = eps, a.gscale = grad_scale;
  a.net = net, a.iter = iter, a.use_ctl = net == 0, a.count = count, a.kl_limit = kl_limit, a.ctl = ctl;
  return launch_peer(c, a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
