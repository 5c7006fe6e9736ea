// mg.cu -- multi-GPU Mul behind the C ABI (reference loop nest src/matrix/mod.rs:965-973 at N GPUs).
//
// Every row of C = A*B needs the same row of A and all of B, so GPU g owns a ROW BLOCK of A and C.  B is not broadcast
// from one owner: rank q owns a COLUMN BLOCK of B (k x n/N, the part it uploads over its own PCIe link or produces
// locally) inside its full-size replica of B, and every rank PULLS the other ranks' column blocks into its replica over
// NVLink with its own kernel (ld.relaxed.sys on the peer-mapped replica, device-side ready/ack flags -- no host
// round trip, no NCCL on the data path).  Column blocks instead of K-panels: C[:, block] = A * B[:, block] needs only that
// block, so every GEMM runs at full depth K with a plain store epilogue (no C read-modify-write), starting with the rank's
// own block while the pulls are in flight.
//
// Two ways in:
//   la_gemm_{f64,f32}_mg        one process, N devices, host operands (what the crate's Mul binds for a large product):
//                               N host threads, contexts cached per (device list, k, n, element size);
//   la_mg_create/handle/connect one process per GPU (torchrun-style launchers): contexts exchange a 256-byte handle
//   + la_gemm_*_mg_rank[_host]  (cudaIpcMemHandle of the replica) through whatever transport the launcher has.
//
// Flags live at the end of each rank's allocation:  ready = epoch of the latest complete part;  ack[q] = last epoch whose
// part rank q has finished pulling.  An owner may rewrite its part only after every ack has reached the previous epoch.
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "la_common.cuh"

namespace la {
namespace {

constexpr int MG_MAX_RANKS = 64;
constexpr int MG_ACK_STRIDE = 32;  // words: one 128-byte line per writer
constexpr size_t MG_FLAG_BYTES = 128 + (size_t)MG_MAX_RANKS * MG_ACK_STRIDE * 4;
constexpr uint32_t MG_MAGIC = 0x6c614d47u;  // "laMG"

struct MgFlags {
  unsigned ready;
  unsigned pad[31];
  unsigned ack[MG_MAX_RANKS * MG_ACK_STRIDE];
};

struct MgHandle {  // LA_MG_HANDLE_BYTES = 256
  uint32_t magic;
  int32_t rank, nranks, device;
  uint64_t pid;
  uint64_t elem, k, n;
  uint64_t raw_base;  // same-process peers use the pointer directly
  cudaIpcMemHandle_t ipc;  // 64 bytes
  unsigned char pad[256 - 4 * 4 - 5 * 8 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(MgHandle) == 256, "handle layout");

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// the owner's part of epoch e is complete in its replica (everything earlier in the stream has finished)
__global__ void mg_publish_kernel(MgFlags* f, unsigned e) {
  __threadfence_system();
  st_release_sys(&f->ready, e);
}
// the owner waits until every peer has pulled epoch e (before it overwrites its part)
__global__ void mg_wait_acks_kernel(const MgFlags* f, int nranks, int self, unsigned e) {
  const int q = threadIdx.x;
  if (q < nranks && q != self) {
    while ((int)(ld_acquire_sys(&f->ack[q * MG_ACK_STRIDE]) - e) < 0) __nanosleep(200);
  }
}
// "I (rank self) have finished reading your part of epoch e" -- written into the OWNER's flag block over NVLink
__global__ void mg_ack_kernel(MgFlags* owner_flags, int self, unsigned e) {
  __threadfence_system();
  st_release_sys(&owner_flags->ack[self * MG_ACK_STRIDE], e);
}
// Pull one column block (rows x width_bytes at byte column col_off of every pitch-byte row) from the owner's replica
// into ours.  All addresses are 16-byte aligned.  Every CTA waits for the owner's ready flag itself (one NVLink
// round trip per poll), then streams with eight 16-byte loads in flight per thread.
constexpr int PULL_THREADS = 512;
constexpr int PULL_UNROLL = 8;
__global__ void __launch_bounds__(PULL_THREADS)
mg_pull_kernel(const char* __restrict__ src, char* __restrict__ dst, size_t pitch, size_t col_off, size_t width_bytes,
               size_t rows, const MgFlags* owner_flags, unsigned e) {
  if (threadIdx.x == 0) {
    while ((int)(ld_acquire_sys(&owner_flags->ready) - e) < 0) __nanosleep(100);
  }
  __syncthreads();
  const size_t chunks_per_row = width_bytes / 16;
  const size_t total = rows * chunks_per_row;
  const size_t stride = (size_t)gridDim.x * PULL_THREADS;
  for (size_t base = (size_t)blockIdx.x * PULL_THREADS + threadIdx.x; base < total; base += stride * PULL_UNROLL) {
    uint4 v[PULL_UNROLL];
    size_t off[PULL_UNROLL];
#pragma unroll
    for (int u = 0; u < PULL_UNROLL; ++u) {
      const size_t c = base + (size_t)u * stride;
      const size_t r = c / chunks_per_row;
      off[u] = r * pitch + col_off + (c - r * chunks_per_row) * 16;
      if (c < total) v[u] = ld_relaxed_sys_v4(src + off[u]);
    }
#pragma unroll
    for (int u = 0; u < PULL_UNROLL; ++u)
      if (base + (size_t)u * stride < total) *reinterpret_cast<uint4*>(dst + off[u]) = v[u];
  }
}

// One launch pulls the column blocks of SEVERAL owners (a whole pass: everything right of our block, or everything left
// of it): the grid is divided evenly over the owners, every CTA waits for its owner's ready flag and streams that block.
// Small blocks (the f32 config: 8 MiB each) are latency-bound -- seven back-to-back (pull, ack) launch pairs cost more than
// the multiply they feed; large ones lose nothing.
struct MgPullList {
  int count;
  const char* src[MG_MAX_RANKS];
  const MgFlags* ready[MG_MAX_RANKS];   // the owner's flag block (its `ready` word is polled)
  unsigned* ack[MG_MAX_RANKS];          // where our ack for that owner goes (its flag block, our slot)
  size_t col_off[MG_MAX_RANKS];
  size_t width_bytes[MG_MAX_RANKS];
};
__global__ void __launch_bounds__(PULL_THREADS)
mg_pull_multi_kernel(const __grid_constant__ MgPullList L, char* __restrict__ dst, size_t pitch, size_t rows, int ctas_per_owner,
                     unsigned e) {
  const int o = blockIdx.x / ctas_per_owner, b = blockIdx.x - o * ctas_per_owner;
  if (threadIdx.x == 0) {
    while ((int)(ld_acquire_sys(&L.ready[o]->ready) - e) < 0) __nanosleep(100);
  }
  __syncthreads();
  const char* __restrict__ src = L.src[o];
  const size_t col_off = L.col_off[o];
  const size_t chunks_per_row = L.width_bytes[o] / 16;
  const size_t total = rows * chunks_per_row;
  const size_t stride = (size_t)ctas_per_owner * PULL_THREADS;
  for (size_t base = (size_t)b * PULL_THREADS + threadIdx.x; base < total; base += stride * PULL_UNROLL) {
    uint4 v[PULL_UNROLL];
    size_t off[PULL_UNROLL];
#pragma unroll
    for (int u = 0; u < PULL_UNROLL; ++u) {
      const size_t c = base + (size_t)u * stride;
      const size_t r = c / chunks_per_row;
      off[u] = r * pitch + col_off + (c - r * chunks_per_row) * 16;
      if (c < total) v[u] = ld_relaxed_sys_v4(src + off[u]);
    }
#pragma unroll
    for (int u = 0; u < PULL_UNROLL; ++u)
      if (base + (size_t)u * stride < total) *reinterpret_cast<uint4*>(dst + off[u]) = v[u];
  }
}
// acks of a whole pass: thread o tells owner o that its block of epoch e has been pulled (runs after the pull kernel)
__global__ void mg_ack_multi_kernel(const __grid_constant__ MgPullList L, unsigned e) {
  __threadfence_system();
  if ((int)threadIdx.x < L.count) st_release_sys(L.ack[threadIdx.x], e);
}

}  // namespace
}  // namespace la

using namespace la;

struct la_mg {
  int rank = 0, nranks = 1, device = 0;
  size_t elem = 8, k = 0, n = 0;
  char* base = nullptr;  // [k * n * elem replica][flags]
  size_t replica_bytes = 0;
  MgFlags* flags = nullptr;
  char* peer_base[MG_MAX_RANKS] = {};
  MgFlags* peer_flags[MG_MAX_RANKS] = {};
  bool peer_ipc[MG_MAX_RANKS] = {};
  size_t col0[MG_MAX_RANKS] = {}, col1[MG_MAX_RANKS] = {};
  unsigned epoch = 0;
  bool connected = false;
  // device copies of the host shards (la_gemm_*_mg_rank_host): owned by the context, not by the calling thread, so that
  // la_mg_reserve can size them while no rank is inside a product (cudaMalloc / cudaFree may wait for the whole device --
  // a rank that allocated BEFORE publishing its block would wait for peers whose kernels are spinning on that block)
  void *shard_a = nullptr, *shard_c = nullptr;
  size_t shard_a_bytes = 0, shard_c_bytes = 0;
  cudaStream_t s_pull = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_pull[2] = {nullptr, nullptr}, ev_up = nullptr, ev_blk[16] = {}, ev_own = nullptr;
  static constexpr int MAX_A_PANELS = 48;
  cudaEvent_t ev_a[MAX_A_PANELS] = {};  // one per uploaded K-panel of the A shard
};

namespace {

struct DevGuard {
  int prev = -1;
  bool active = false;
  int enter(int device) {
    LA_CUDA_TRY(cudaGetDevice(&prev));
    if (prev != device) {
      LA_CUDA_TRY(cudaSetDevice(device));
      active = true;
    }
    return LA_OK;
  }
  ~DevGuard() {
    if (active) cudaSetDevice(prev);
  }
};

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// LA_MG_TRACE=1: stage markers of every rank on stderr (diagnostic)
bool mg_trace_on() {
  static const bool v = getenv("LA_MG_TRACE") && atoi(getenv("LA_MG_TRACE")) != 0;
  return v;
}
#define MG_TRACE(c, ...)                                                     \
  do {                                                                       \
    if (mg_trace_on()) {                                                     \
      fprintf(stderr, "[mg rank %d epoch %u] ", (c)->rank, (c)->epoch);      \
      fprintf(stderr, __VA_ARGS__);                                          \
      fprintf(stderr, "\n");                                                 \
      fflush(stderr);                                                        \
    }                                                                        \
  } while (0)

// Block partition of `total` into `parts` pieces whose boundaries are multiples of `align` where possible.
void block_range(size_t total, int parts, int idx, size_t align, size_t* b0, size_t* b1) {
  const size_t units = (total + align - 1) / align;
  const size_t per = units / (size_t)parts, extra = units % (size_t)parts;
  const size_t u0 = per * (size_t)idx + ((size_t)idx < extra ? (size_t)idx : extra);
  const size_t u1 = u0 + per + ((size_t)idx < extra ? 1 : 0);
  *b0 = u0 * align < total ? u0 * align : total;
  *b1 = u1 * align < total ? u1 * align : total;
}

int mg_pull_ctas() {
  static const int v = getenv("LA_MG_PULL_CTAS") ? atoi(getenv("LA_MG_PULL_CTAS")) : 32;
  return v < 1 ? 1 : v;
}
// LA_MG_PULL=ce: move the blocks with the copy engines (cudaMemcpy2DAsync on the peer-mapped pointer) behind the same
// device-side flags; default: the SM pull kernel above.
bool mg_pull_with_ce() {
  static const bool v = getenv("LA_MG_PULL") && !strcmp(getenv("LA_MG_PULL"), "ce");
  return v;
}

// Queues on ctx->s_pull: every other rank's column block of epoch e into our replica (blocks right of ours first, then
// the ones left of it), an ack per block, and the two events the GEMMs of those column ranges wait for.
int mg_queue_pulls(la_mg* c, unsigned e) {
  const size_t pitch = c->n * c->elem;
  for (int pass = 0; pass < 2; ++pass) {
    const int q0 = pass == 0 ? c->rank + 1 : 0, q1 = pass == 0 ? c->nranks : c->rank;
    if (mg_pull_with_ce()) {
      for (int q = q0; q < q1; ++q) {
        const size_t off = c->col0[q] * c->elem, wb = (c->col1[q] - c->col0[q]) * c->elem;
        if (wb == 0) continue;
        // the flag wait still happens on the device: a one-CTA pull of zero rows
        mg_pull_kernel<<<1, PULL_THREADS, 0, c->s_pull>>>(c->peer_base[q], c->base, pitch, off, wb, 0, c->peer_flags[q], e);
        LA_CUDA_TRY(cudaMemcpy2DAsync(c->base + off, pitch, c->peer_base[q] + off, pitch, wb, c->k, cudaMemcpyDeviceToDevice,
                                      c->s_pull));
        mg_ack_kernel<<<1, 1, 0, c->s_pull>>>(c->peer_flags[q], c->rank, e);
        LA_CUDA_TRY(cudaGetLastError());
      }
    } else {
      MgPullList L;
      L.count = 0;
      for (int q = q0; q < q1; ++q) {
        const size_t wb = (c->col1[q] - c->col0[q]) * c->elem;
        if (wb == 0) continue;
        const int o = L.count++;
        L.src[o] = c->peer_base[q];
        L.ready[o] = c->peer_flags[q];
        L.ack[o] = &c->peer_flags[q]->ack[c->rank * MG_ACK_STRIDE];
        L.col_off[o] = c->col0[q] * c->elem;
        L.width_bytes[o] = wb;
      }
      if (L.count > 0) {
        // mg_pull_ctas() CTAs per owner, at most twice that in total (pull CTAs take SMs from the GEMM that runs beside them)
        int per = mg_pull_ctas();
        if (per * L.count > 2 * mg_pull_ctas()) per = (2 * mg_pull_ctas() + L.count - 1) / L.count;
        mg_pull_multi_kernel<<<per * L.count, PULL_THREADS, 0, c->s_pull>>>(L, c->base, pitch, c->k, per, e);
        mg_ack_multi_kernel<<<1, MG_MAX_RANKS, 0, c->s_pull>>>(L, e);
        LA_CUDA_TRY(cudaGetLastError());
      }
    }
    LA_CUDA_TRY(cudaEventRecord(c->ev_pull[pass], c->s_pull));
  }
  return LA_OK;
}

int mg_reserve(la_mg* c, size_t m_local) {
  LA_REQUIRE(c && m_local > 0, "la_mg_reserve: bad arguments");
  DevGuard g;
  LA_TRY(g.enter(c->device));
  const size_t need_a = m_local * c->k * c->elem, need_c = m_local * c->n * c->elem;
  if (c->shard_a_bytes < need_a) {
    if (c->shard_a) LA_CUDA_TRY(cudaFree(c->shard_a));
    c->shard_a = nullptr;
    c->shard_a_bytes = 0;
    LA_CUDA_TRY(cudaMalloc(&c->shard_a, need_a));
    c->shard_a_bytes = need_a;
  }
  if (c->shard_c_bytes < need_c) {
    if (c->shard_c) LA_CUDA_TRY(cudaFree(c->shard_c));
    c->shard_c = nullptr;
    c->shard_c_bytes = 0;
    LA_CUDA_TRY(cudaMalloc(&c->shard_c, need_c));
    c->shard_c_bytes = need_c;
  }
  return LA_OK;
}

template <typename T>
int mg_check(const la_mg* c, const char* who) {
  LA_REQUIRE(c != nullptr, "%s: null context", who);
  LA_REQUIRE(c->connected, "%s: la_mg_connect has not been called on this context", who);
  LA_REQUIRE(c->elem == sizeof(T), "%s: context was created for %zu-byte elements", who, c->elem);
  return LA_OK;
}

// Device-resident shard: the rank's own column block is already in its replica (stream-ordered before this call).
template <typename T>
int mg_rank_dev(la_mg* c, const T* A, size_t lda, T* C, size_t ldc, size_t m_local, cudaStream_t st) {
  LA_TRY(mg_check<T>(c, "la_gemm_mg_rank"));
  LA_REQUIRE(A && C && m_local > 0 && lda >= c->k && ldc >= c->n, "la_gemm_mg_rank: bad shard arguments");
  DevGuard g;
  LA_TRY(g.enter(c->device));
  CallScope scope(c->device, st);
  const unsigned e = ++c->epoch;
  T* B = reinterpret_cast<T*>(c->base);
  const size_t n = c->n, k = c->k;
  // our pulls may overwrite blocks that the previous call's GEMMs (earlier on `st`) still read
  LA_CUDA_TRY(cudaEventRecord(c->ev_begin, st));
  LA_CUDA_TRY(cudaStreamWaitEvent(c->s_pull, c->ev_begin, 0));
  mg_publish_kernel<<<1, 1, 0, st>>>(c->flags, e);
  LA_CUDA_TRY(cudaGetLastError());
  LA_TRY(mg_queue_pulls(c, e));
  const size_t o0 = c->col0[c->rank], o1 = c->col1[c->rank];
  if (o1 > o0) LA_TRY(gemm_dev<T>(A, lda, B + o0, n, C + o0, ldc, m_local, k, o1 - o0, LA_GEMM_ASSIGN, st));
  if (o1 < n) {  // the blocks right of ours
    LA_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_pull[0], 0));
    LA_TRY(gemm_dev<T>(A, lda, B + o1, n, C + o1, ldc, m_local, k, n - o1, LA_GEMM_ASSIGN, st));
  }
  if (o0 > 0) {  // the blocks left of ours
    LA_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_pull[1], 0));
    LA_TRY(gemm_dev<T>(A, lda, B, n, C, ldc, m_local, k, o0, LA_GEMM_ASSIGN, st));
  }
  // the caller's stream also covers the pull stream's tail (acks are queued behind the pulls)
  LA_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_pull[1], 0));
  return LA_OK;
}

// Host shard: A rows [m_local x k] (tight), this rank's column block of B (k rows, leading dimension ldb_host elements),
// C rows out [m_local x n] (tight).  Synchronous.  Pipeline: own block in K-panels under its own upload, the other column
// ranges at full depth while the pulls land, the last range in shrinking row blocks whose finished rows of C go down
// while the next block multiplies.
template <typename T>
int mg_rank_host(la_mg* c, const T* A, const T* Bblk, size_t ldb_host, T* C, size_t m_local) {
  LA_TRY(mg_check<T>(c, "la_gemm_mg_rank_host"));
  LA_REQUIRE(A && Bblk && C && m_local > 0, "la_gemm_mg_rank_host: bad shard arguments");
  DevGuard g;
  LA_TRY(g.enter(c->device));
  const size_t n = c->n, k = c->k, es = sizeof(T);
  const size_t o0 = c->col0[c->rank], o1 = c->col1[c->rank], ow = o1 - o0;
  LA_REQUIRE(ldb_host >= ow, "la_gemm_mg_rank_host: ldb smaller than the column block");
  MG_TRACE(c, "host call: m_local=%zu", m_local);
  LA_TRY(mg_reserve(c, m_local));  // no-op after la_mg_reserve / an earlier call with the same shard height
  MG_TRACE(c, "shard buffers ready");
  T* Ad = (T*)c->shard_a;
  T* Cd = (T*)c->shard_c;
  T* B = reinterpret_cast<T*>(c->base);
  cudaStream_t st = cudaStreamPerThread, up = c->s_h2d, down = c->s_d2h;
  CallScope scope(c->device, st);
  const unsigned e = ++c->epoch;

  // nobody may still be reading the previous epoch's part when the upload overwrites it
  if (e > 1) mg_wait_acks_kernel<<<1, MG_MAX_RANKS, 0, up>>>(c->flags, c->nranks, c->rank, e - 1);
  LA_CUDA_TRY(cudaGetLastError());
  LA_CUDA_TRY(cudaEventRecord(c->ev_begin, st));
  LA_CUDA_TRY(cudaStreamWaitEvent(c->s_pull, c->ev_begin, 0));
  LA_CUDA_TRY(cudaStreamWaitEvent(up, c->ev_begin, 0));

  // ---- own column block: K-panels (first one narrow), B part first so that it can be published early ----
  const bool deep = k >= 2048 && ow > 0;
  if (ow > 0)
    LA_CUDA_TRY(cudaMemcpy2DAsync(B + o0, n * es, Bblk, ldb_host * es, ow * es, k, cudaMemcpyHostToDevice, up));
  MG_TRACE(c, "own block upload queued");
  mg_publish_kernel<<<1, 1, 0, up>>>(c->flags, e);
  LA_CUDA_TRY(cudaGetLastError());
  // The A shard goes up in K-panels (first one narrow) so that the multiply of the own block starts under its own upload.
  // EVERY upload is queued before the pulls: a copy from pageable memory is staged by the driver and would otherwise
  // wait behind this rank's own pull kernels, which spin until the peers have published.
  struct Panel {
    size_t k0, w;
  };
  Panel pan[la_mg::MAX_A_PANELS];
  int npan = 0;
  if (deep) {
    const size_t kp = 2048;
    for (size_t k0 = 0; k0 < k;) {
      size_t w = (npan == 0) ? 512 : ((npan == 1) ? kp - 512 : kp);
      if (w > k - k0 || npan == la_mg::MAX_A_PANELS - 1) w = k - k0;
      pan[npan++] = {k0, w};
      k0 += w;
    }
  } else {
    pan[npan++] = {0, k};
  }
  for (int p = 0; p < npan; ++p) {
    if (npan == 1)
      LA_CUDA_TRY(cudaMemcpyAsync(Ad, A, m_local * k * es, cudaMemcpyHostToDevice, up));
    else
      LA_CUDA_TRY(cudaMemcpy2DAsync(Ad + pan[p].k0, k * es, A + pan[p].k0, k * es, pan[p].w * es, m_local,
                                    cudaMemcpyHostToDevice, up));
    LA_CUDA_TRY(cudaEventRecord(c->ev_a[p], up));
  }
  MG_TRACE(c, "A shard uploads queued");
  LA_TRY(mg_queue_pulls(c, e));
  MG_TRACE(c, "publish + pulls queued");
  for (int p = 0; p < npan; ++p) {
    LA_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_a[p], 0));
    if (ow > 0)
      LA_TRY(gemm_dev<T>(Ad + pan[p].k0, k, B + pan[p].k0 * n + o0, n, Cd + o0, n, m_local, pan[p].w, ow,
                         p == 0 ? LA_GEMM_ASSIGN : LA_GEMM_ADD, st));
  }
  // ---- the other column ranges; the LAST non-empty one runs in row blocks with the download of C behind it ----
  struct Range {
    size_t c0, c1;
    cudaEvent_t ev;
  };
  Range rg[2];
  int nrg = 0;
  if (o1 < n) rg[nrg++] = {o1, n, c->ev_pull[0]};
  if (o0 > 0) rg[nrg++] = {0, o0, c->ev_pull[1]};
  for (int i = 0; i + 1 < nrg; ++i) {
    LA_CUDA_TRY(cudaStreamWaitEvent(st, rg[i].ev, 0));
    LA_TRY(gemm_dev<T>(Ad, k, B + rg[i].c0, n, Cd + rg[i].c0, n, m_local, k, rg[i].c1 - rg[i].c0, LA_GEMM_ASSIGN, st));
  }
  // row blocks: 1/4, 1/4, 1/4, 1/8, 1/16, 1/16 of the rows (multiples of 128) when the shard is tall enough
  size_t blk[8];
  int nblk = 0;
  if (m_local >= 16 * 128) {
    const size_t q = round_up(m_local / 4, 128), h = round_up(m_local / 8, 128), x = round_up(m_local / 16, 128);
    const size_t want[6] = {q, q, q, h, x, x};
    size_t left = m_local;
    for (int i = 0; i < 6 && left > 0; ++i) {
      const size_t take = (i == 5 || want[i] > left) ? left : want[i];
      blk[nblk++] = take;
      left -= take;
    }
  } else {
    blk[nblk++] = m_local;
  }
  if (nrg > 0) LA_CUDA_TRY(cudaStreamWaitEvent(st, rg[nrg - 1].ev, 0));
  size_t r0 = 0;
  for (int b = 0; b < nblk; r0 += blk[b], ++b) {
    if (nrg > 0) {
      const Range& R = rg[nrg - 1];
      LA_TRY(gemm_dev<T>(Ad + r0 * k, k, B + R.c0, n, Cd + r0 * n + R.c0, n, blk[b], k, R.c1 - R.c0, LA_GEMM_ASSIGN, st));
    }
    LA_CUDA_TRY(cudaEventRecord(c->ev_blk[b], st));
    LA_CUDA_TRY(cudaStreamWaitEvent(down, c->ev_blk[b], 0));
    LA_CUDA_TRY(cudaMemcpyAsync(C + r0 * n, Cd + r0 * n, blk[b] * n * es, cudaMemcpyDeviceToHost, down));
  }
  MG_TRACE(c, "all work queued, synchronising");
  LA_CUDA_TRY(cudaStreamSynchronize(down));
  MG_TRACE(c, "downloads done");
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  LA_CUDA_TRY(cudaStreamSynchronize(c->s_pull));  // our acks are out: the peers may move on
  MG_TRACE(c, "call complete");
  return LA_OK;
}

int mg_create(int rank, int nranks, int device, size_t elem, size_t k, size_t n, la_mg** out) {
  LA_REQUIRE(out, "la_mg_create: null output");
  *out = nullptr;
  LA_REQUIRE(nranks >= 1 && nranks <= MG_MAX_RANKS && rank >= 0 && rank < nranks, "la_mg_create: bad rank %d of %d", rank,
             nranks);
  LA_REQUIRE(elem == 4 || elem == 8, "la_mg_create: element size must be 4 or 8");
  LA_REQUIRE(k > 0 && n > 0, "la_mg_create: zero dimension");
  LA_REQUIRE((n * elem) % 16 == 0, "la_mg_create: a row of B must be a multiple of 16 bytes (n = %zu)", n);
  const DeviceCtx* ctx;
  LA_TRY(device_ctx(device, &ctx));
  DevGuard g;
  LA_TRY(g.enter(device));
  la_mg* c = new la_mg();
  c->rank = rank;
  c->nranks = nranks;
  c->device = device;
  c->elem = elem;
  c->k = k;
  c->n = n;
  c->replica_bytes = round_up(k * n * elem, 256);
  cudaError_t err = cudaMalloc(&c->base, c->replica_bytes + MG_FLAG_BYTES);
  if (err != cudaSuccess) {
    const size_t want = c->replica_bytes;
    delete c;
    cudaGetLastError();
    return fail(LA_ERR_NOMEM, "la_mg_create: cudaMalloc of the %zu-byte replica of B failed: %s", want,
                cudaGetErrorString(err));
  }
  c->flags = reinterpret_cast<MgFlags*>(c->base + c->replica_bytes);
  LA_CUDA_TRY(cudaMemset(c->flags, 0, MG_FLAG_BYTES));
  // column blocks: multiples of 256 columns (the f32 tile width; two f64 tiles) where n allows it
  const size_t align = (n / (size_t)nranks >= 256) ? 256 : 16 / elem;
  for (int q = 0; q < nranks; ++q) block_range(n, nranks, q, align, &c->col0[q], &c->col1[q]);
  LA_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_pull, cudaStreamNonBlocking));
  LA_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  LA_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  LA_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_begin, cudaEventDisableTiming));
  LA_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pull[0], cudaEventDisableTiming));
  LA_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pull[1], cudaEventDisableTiming));
  LA_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
  LA_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_own, cudaEventDisableTiming));
  for (int i = 0; i < 16; ++i) LA_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_blk[i], cudaEventDisableTiming));
  for (int i = 0; i < la_mg::MAX_A_PANELS; ++i) LA_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_a[i], cudaEventDisableTiming));
  c->peer_base[rank] = c->base;
  c->peer_flags[rank] = c->flags;
  if (nranks == 1) c->connected = true;
  *out = c;
  return LA_OK;
}

int mg_handle(const la_mg* c, void* out) {
  LA_REQUIRE(c && out, "la_mg_handle: null pointer");
  DevGuard g;
  LA_TRY(g.enter(c->device));
  MgHandle h;
  memset(&h, 0, sizeof(h));
  h.magic = MG_MAGIC;
  h.rank = c->rank;
  h.nranks = c->nranks;
  h.device = c->device;
  h.pid = (uint64_t)getpid();
  h.elem = c->elem;
  h.k = c->k;
  h.n = c->n;
  h.raw_base = (uint64_t)(uintptr_t)c->base;
  LA_CUDA_TRY(cudaIpcGetMemHandle(&h.ipc, c->base));
  memcpy(out, &h, sizeof(h));
  return LA_OK;
}

int mg_connect(la_mg* c, const void* handles) {
  LA_REQUIRE(c && handles, "la_mg_connect: null pointer");
  LA_REQUIRE(!c->connected || c->nranks == 1, "la_mg_connect: already connected");
  DevGuard g;
  LA_TRY(g.enter(c->device));
  const MgHandle* H = reinterpret_cast<const MgHandle*>(handles);
  for (int q = 0; q < c->nranks; ++q) {
    MgHandle h;
    memcpy(&h, &H[q], sizeof(h));
    LA_REQUIRE(h.magic == MG_MAGIC && h.rank == q && h.nranks == c->nranks, "la_mg_connect: handle %d is not rank %d of %d", q,
               q, c->nranks);
    LA_REQUIRE(h.elem == c->elem && h.k == c->k && h.n == c->n, "la_mg_connect: rank %d was created for a different B", q);
    if (q == c->rank) continue;
    char* p = nullptr;
    if (h.pid == (uint64_t)getpid()) {
      if (h.device != c->device) {
        int can = 0;
        LA_CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->device, h.device));
        if (!can) return fail(LA_ERR_UNSUPPORTED, "la_mg_connect: device %d cannot access device %d", c->device, h.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(LA_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", h.device, cudaGetErrorString(e));
        cudaGetLastError();
      }
      p = reinterpret_cast<char*>((uintptr_t)h.raw_base);
    } else {
      void* vp = nullptr;
      LA_CUDA_TRY(cudaIpcOpenMemHandle(&vp, h.ipc, cudaIpcMemLazyEnablePeerAccess));
      p = (char*)vp;
      c->peer_ipc[q] = true;
    }
    c->peer_base[q] = p;
    c->peer_flags[q] = reinterpret_cast<MgFlags*>(p + c->replica_bytes);
  }
  // Load every kernel a product launches, and run the copy paths once, while no rank can be waiting on this one
  // (lazy kernel loading and the driver's staging of 2-D copies may wait for the device; a peer's pull kernel that spins
  // on OUR publish would make that wait circular when ranks share a device).
  {
    cudaFuncAttributes fa;
    LA_CUDA_TRY(cudaFuncGetAttributes(&fa, mg_publish_kernel));
    LA_CUDA_TRY(cudaFuncGetAttributes(&fa, mg_wait_acks_kernel));
    LA_CUDA_TRY(cudaFuncGetAttributes(&fa, mg_ack_kernel));
    LA_CUDA_TRY(cudaFuncGetAttributes(&fa, mg_pull_kernel));
    LA_CUDA_TRY(cudaFuncGetAttributes(&fa, mg_pull_multi_kernel));
    LA_CUDA_TRY(cudaFuncGetAttributes(&fa, mg_ack_multi_kernel));
    LA_TRY(gemm_f64_preload());
    LA_TRY(gemm_f32_preload());
    LA_TRY(gemm_simt_preload());
    unsigned char warm[64] = {0};
    char* pad = reinterpret_cast<char*>(c->flags) + 16;  // inside MgFlags::pad: never read by the protocol
    LA_CUDA_TRY(cudaMemcpy2DAsync(pad, 32, warm, 16, 16, 2, cudaMemcpyHostToDevice, c->s_h2d));
    LA_CUDA_TRY(cudaMemcpy2DAsync(warm + 32, 16, pad, 32, 16, 2, cudaMemcpyDeviceToHost, c->s_d2h));
    LA_CUDA_TRY(cudaStreamSynchronize(c->s_h2d));
    LA_CUDA_TRY(cudaStreamSynchronize(c->s_d2h));
  }
  c->connected = true;
  return LA_OK;
}

int mg_destroy(la_mg* c) {
  if (!c) return LA_OK;
  DevGuard g;
  LA_TRY(g.enter(c->device));
  cudaDeviceSynchronize();
  for (int q = 0; q < c->nranks; ++q)
    if (c->peer_ipc[q] && c->peer_base[q]) cudaIpcCloseMemHandle(c->peer_base[q]);
  if (c->s_pull) cudaStreamDestroy(c->s_pull);
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  cudaEventDestroy(c->ev_begin);
  cudaEventDestroy(c->ev_pull[0]);
  cudaEventDestroy(c->ev_pull[1]);
  cudaEventDestroy(c->ev_up);
  cudaEventDestroy(c->ev_own);
  for (int i = 0; i < 16; ++i) cudaEventDestroy(c->ev_blk[i]);
  for (int i = 0; i < la_mg::MAX_A_PANELS; ++i) cudaEventDestroy(c->ev_a[i]);
  cudaFree(c->base);
  if (c->shard_a) cudaFree(c->shard_a);
  if (c->shard_c) cudaFree(c->shard_c);
  cudaGetLastError();
  delete c;
  return LA_OK;
}

// ---- single-process form: contexts cached per (element size, k, n, device list), one persistent host thread per device
//      (the per-thread scratch pools of the GEMM kernels live as long as the group) ----
struct MgWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, done = true, quit = false;
  MgWorker() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv.wait(lk, [this] { return has_job || quit; });
        if (quit) return;
        std::function<void()> j = std::move(job);
        has_job = false;
        lk.unlock();
        j();
        lk.lock();
        done = true;
        cv.notify_all();
      }
    });
  }
  void submit(std::function<void()> j) {
    std::lock_guard<std::mutex> lk(m);
    job = std::move(j);
    has_job = true;
    done = false;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return done; });
  }
  ~MgWorker() {
    {
      std::lock_guard<std::mutex> lk(m);
      quit = true;
      cv.notify_all();
    }
    th.join();
  }
};
struct MgGroup {
  std::vector<la_mg*> ctx;
  std::vector<MgWorker*> workers;
  std::mutex busy;  // one product at a time per group (the flags are per-context epochs)
  ~MgGroup() {
    for (MgWorker* w : workers) delete w;
    for (la_mg* c : ctx) mg_destroy(c);
  }
};
std::mutex g_groups_mutex;
std::map<std::string, MgGroup*> g_groups;

int mg_group_get(int ngpus, const int* devices, size_t elem, size_t k, size_t n, MgGroup** out) {
  const std::string devkey((const char*)devices, sizeof(int) * (size_t)ngpus);
  const std::string key = devkey + "|" + std::to_string(elem) + "|" + std::to_string(k) + "|" + std::to_string(n);
  std::lock_guard<std::mutex> lock(g_groups_mutex);
  auto it = g_groups.find(key);
  if (it != g_groups.end()) {
    *out = it->second;
    return LA_OK;
  }
  // a new shape replaces the idle cached groups of the same device list (each holds a full replica of B per device)
  for (auto jt = g_groups.begin(); jt != g_groups.end();) {
    if (jt->first.compare(0, devkey.size() + 1, devkey + "|") == 0 && jt->second->busy.try_lock()) {
      jt->second->busy.unlock();
      delete jt->second;
      jt = g_groups.erase(jt);
    } else {
      ++jt;
    }
  }
  MgGroup* G = new MgGroup();
  std::vector<MgHandle> handles((size_t)ngpus);
  int status = LA_OK;
  for (int r = 0; r < ngpus && status == LA_OK; ++r) {
    la_mg* c = nullptr;
    status = mg_create(r, ngpus, devices[r], elem, k, n, &c);
    if (status == LA_OK) {
      G->ctx.push_back(c);
      status = mg_handle(c, &handles[(size_t)r]);
    }
  }
  for (int r = 0; r < ngpus && status == LA_OK; ++r) status = mg_connect(G->ctx[(size_t)r], handles.data());
  if (status != LA_OK) {
    const std::string keep = error_text();
    delete G;
    set_error("%s", keep.c_str());
    return status;
  }
  for (int r = 0; r < ngpus; ++r) G->workers.push_back(new MgWorker());
  g_groups[key] = G;
  *out = G;
  return LA_OK;
}

template <typename T>
int gemm_host_single(const T* A, const T* B, T* C, size_t m, size_t k, size_t n);
template <>
int gemm_host_single<double>(const double* A, const double* B, double* C, size_t m, size_t k, size_t n) {
  return la_gemm_f64_host(A, B, C, m, k, n);
}
template <>
int gemm_host_single<float>(const float* A, const float* B, float* C, size_t m, size_t k, size_t n) {
  return la_gemm_f32_host(A, B, C, m, k, n);
}

template <typename T>
int gemm_mg(int ngpus, const int* devices, const T* A, const T* B, T* C, size_t m, size_t k, size_t n) {
  LA_REQUIRE(devices && A && B && C, "la_gemm_mg: null pointer");
  LA_REQUIRE(ngpus >= 1 && ngpus <= MG_MAX_RANKS, "la_gemm_mg: bad device count %d", ngpus);
  LA_REQUIRE(m > 0 && k > 0 && n > 0, "la_gemm_mg: zero dimension (m=%zu k=%zu n=%zu)", m, k, n);
  for (int i = 0; i < ngpus; ++i)
    for (int j = 0; j < i; ++j) LA_REQUIRE(devices[i] != devices[j], "la_gemm_mg: device %d listed twice", devices[i]);
  // every rank gets at least one 128-row tile band; a product too small (or too oddly shaped) to share runs on devices[0]
  const size_t bands = (m + 127) / 128;
  const int eff = (size_t)ngpus < bands ? ngpus : (int)bands;
  if (eff == 1 || (n * sizeof(T)) % 16 != 0) {
    DevGuard g;
    LA_TRY(g.enter(devices[0]));
    return gemm_host_single<T>(A, B, C, m, k, n);
  }
  MgGroup* G = nullptr;
  LA_TRY(mg_group_get(eff, devices, sizeof(T), k, n, &G));
  std::lock_guard<std::mutex> lock(G->busy);
  std::vector<int> status((size_t)eff, LA_OK);
  std::vector<std::string> text((size_t)eff);
  // phase 1: every rank sizes its shard buffers; nobody publishes or pulls before all allocations are done
  for (int r = 0; r < eff; ++r) {
    G->workers[(size_t)r]->submit([&, r] {
      size_t r0, r1;
      block_range(m, eff, r, 128, &r0, &r1);
      const int s = mg_reserve(G->ctx[(size_t)r], r1 - r0);
      status[(size_t)r] = s;
      if (s != LA_OK) text[(size_t)r] = error_text();
    });
  }
  for (int r = 0; r < eff; ++r) G->workers[(size_t)r]->wait();
  for (int r = 0; r < eff; ++r)
    if (status[(size_t)r] != LA_OK) return fail(status[(size_t)r], "la_gemm_mg (rank %d): %s", r, text[(size_t)r].c_str());
  // phase 2: the product
  for (int r = 0; r < eff; ++r) {
    G->workers[(size_t)r]->submit([&, r] {
      la_mg* c = G->ctx[(size_t)r];
      size_t r0, r1;
      block_range(m, eff, r, 128, &r0, &r1);
      int s = cudaSetDevice(c->device) == cudaSuccess ? LA_OK : fail(LA_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
      if (s == LA_OK) s = mg_rank_host<T>(c, A + r0 * k, B + c->col0[r], n, C + r0 * n, r1 - r0);
      status[(size_t)r] = s;
      if (s != LA_OK) text[(size_t)r] = error_text();
    });
  }
  for (int r = 0; r < eff; ++r) G->workers[(size_t)r]->wait();
  for (int r = 0; r < eff; ++r)
    if (status[(size_t)r] != LA_OK) return fail(status[(size_t)r], "la_gemm_mg (rank %d): %s", r, text[(size_t)r].c_str());
  return LA_OK;
}

}  // namespace

extern "C" {

int la_mg_shard(int nranks, int rank, size_t m, size_t n, size_t elem_bytes, size_t* row0, size_t* row1, size_t* col0,
                size_t* col1) {
  LA_REQUIRE(nranks >= 1 && nranks <= MG_MAX_RANKS && rank >= 0 && rank < nranks, "la_mg_shard: bad rank %d of %d", rank, nranks);
  LA_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "la_mg_shard: element size must be 4 or 8");
  LA_REQUIRE(row0 && row1 && col0 && col1, "la_mg_shard: null output");
  block_range(m, nranks, rank, 128, row0, row1);
  const size_t align = (n / (size_t)nranks >= 256) ? 256 : 16 / elem_bytes;
  block_range(n, nranks, rank, align, col0, col1);
  return LA_OK;
}
int la_mg_create(int rank, int nranks, int device, size_t elem_bytes, size_t k, size_t n, la_mg** out) {
  return mg_create(rank, nranks, device, elem_bytes, k, n, out);
}
int la_mg_handle(const la_mg* ctx, void* handle_out) { return mg_handle(ctx, handle_out); }
int la_mg_connect(la_mg* ctx, const void* handles) { return mg_connect(ctx, handles); }
int la_mg_destroy(la_mg* ctx) { return mg_destroy(ctx); }
int la_mg_b_block(const la_mg* ctx, void** block_dev, size_t* ldb, size_t* col0, size_t* col1) {
  LA_REQUIRE(ctx && block_dev && ldb && col0 && col1, "la_mg_b_block: null pointer");
  *block_dev = ctx->base + ctx->col0[ctx->rank] * ctx->elem;
  *ldb = ctx->n;
  *col0 = ctx->col0[ctx->rank];
  *col1 = ctx->col1[ctx->rank];
  return LA_OK;
}
int la_mg_reserve(la_mg* ctx, size_t m_local) { return mg_reserve(ctx, m_local); }
int la_mg_quiesce(la_mg* ctx, void* cuda_stream) {
  LA_REQUIRE(ctx && ctx->connected, "la_mg_quiesce: context not connected");
  DevGuard g;
  LA_TRY(g.enter(ctx->device));
  if (ctx->epoch > 0 && ctx->nranks > 1)
    mg_wait_acks_kernel<<<1, MG_MAX_RANKS, 0, resolve_stream(cuda_stream)>>>(ctx->flags, ctx->nranks, ctx->rank, ctx->epoch);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
int la_gemm_f64_mg_rank(la_mg* ctx, const double* A_shard, size_t lda, double* C_shard, size_t ldc, size_t m_local,
                        void* cuda_stream) {
  return mg_rank_dev<double>(ctx, A_shard, lda, C_shard, ldc, m_local, resolve_stream(cuda_stream));
}
int la_gemm_f32_mg_rank(la_mg* ctx, const float* A_shard, size_t lda, float* C_shard, size_t ldc, size_t m_local,
                        void* cuda_stream) {
  return mg_rank_dev<float>(ctx, A_shard, lda, C_shard, ldc, m_local, resolve_stream(cuda_stream));
}
int la_gemm_f64_mg_rank_host(la_mg* ctx, const double* A_shard, const double* B_block, size_t ldb, double* C_shard,
                             size_t m_local) {
  return mg_rank_host<double>(ctx, A_shard, B_block, ldb, C_shard, m_local);
}
int la_gemm_f32_mg_rank_host(la_mg* ctx, const float* A_shard, const float* B_block, size_t ldb, float* C_shard,
                             size_t m_local) {
  return mg_rank_host<float>(ctx, A_shard, B_block, ldb, C_shard, m_local);
}
int la_gemm_f64_mg(int ngpus, const int* devices, const double* A, const double* B, double* C, size_t m, size_t k, size_t n) {
  return gemm_mg<double>(ngpus, devices, A, B, C, m, k, n);
}
int la_gemm_f32_mg(int ngpus, const int* devices, const float* A, const float* B, float* C, size_t m, size_t k, size_t n) {
  return gemm_mg<float>(ngpus, devices, A, B, C, m, k, n);
}

}  // extern "C"
