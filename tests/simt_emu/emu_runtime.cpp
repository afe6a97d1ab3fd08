// tests/simt_emu/emu_runtime.cpp -- TEST INFRASTRUCTURE ONLY: SIMT interpreter + CUDA runtime stand-ins (see
// include/cuda_runtime.h).  One OS thread; the CUDA threads of one block are ucontext fibers that run until they reach
// a scheduling point (block / warp / named barrier, shuffle) or return; blocks run one after the other.
#include <dlfcn.h>
#include <deque>
#include <cuda_runtime.h>
#include <mif_tma.cuh>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <ucontext.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <string>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace {

enum State { RUNNABLE, WAIT_BLOCK, WAIT_WARP, WAIT_NAMED, DONE };
constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Fiber {
  ucontext_t ctx;
  State state = DONE;
  int named_id = 0, named_count = 0;
  char *stack = nullptr;
};

Fiber g_fibers[kMaxThreads];
ucontext_t g_scheduler;
int g_current = -1, g_nthreads = 0;
const std::function<void()> *g_body = nullptr;
std::vector<unsigned char> g_dynamic_smem;
uint64_t g_shuffle[kMaxThreads / 32][2][32];
int g_parity[kMaxThreads];
uint3 g_tid[kMaxThreads];

void trampoline() {
  (*g_body)();
  g_fibers[g_current].state = DONE;
  swapcontext(&g_fibers[g_current].ctx, &g_scheduler);
}

void yield(State why) {
  Fiber &f = g_fibers[g_current];
  f.state = why;
  swapcontext(&f.ctx, &g_scheduler);
  threadIdx = g_tid[g_current];
}

void run_block() {
  for (int t = 0; t < g_nthreads; t++) {
    Fiber &f = g_fibers[t];
    if (!f.stack) f.stack = static_cast<char *>(aligned_alloc(64, kStackBytes));
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, trampoline, 0);
    f.state = RUNNABLE;
    g_parity[t] = 0;
  }
  for (;;) {
    bool ran = false;
    for (int t = 0; t < g_nthreads; t++) {
      if (g_fibers[t].state != RUNNABLE) continue;
      g_current = t;
      threadIdx = g_tid[t];
      swapcontext(&g_scheduler, &g_fibers[t].ctx);
      ran = true;
    }
    // release the barriers whose participants have all arrived (threads that returned do not take part)
    int live = 0, at_block = 0;
    for (int t = 0; t < g_nthreads; t++) {
      if (g_fibers[t].state != DONE) live++;
      if (g_fibers[t].state == WAIT_BLOCK) at_block++;
    }
    if (live == 0) {
      emu::tma_block_end();
      return;
    }
    bool released = false;
    if (at_block == live) {
      for (int t = 0; t < g_nthreads; t++)
        if (g_fibers[t].state == WAIT_BLOCK) g_fibers[t].state = RUNNABLE;
      released = true;
    }
    for (int w = 0; w * 32 < g_nthreads; w++) {
      int wl = 0, ww = 0;
      for (int t = w * 32; t < g_nthreads && t < w * 32 + 32; t++) {
        if (g_fibers[t].state != DONE) wl++;
        if (g_fibers[t].state == WAIT_WARP) ww++;
      }
      if (wl > 0 && ww == wl) {
        for (int t = w * 32; t < g_nthreads && t < w * 32 + 32; t++)
          if (g_fibers[t].state == WAIT_WARP) g_fibers[t].state = RUNNABLE;
        released = true;
      }
    }
    for (int id = 0; id < 16; id++) {
      int waiting = 0, expected = 0;
      for (int t = 0; t < g_nthreads; t++)
        if (g_fibers[t].state == WAIT_NAMED && g_fibers[t].named_id == id) {
          waiting++;
          expected = g_fibers[t].named_count;
        }
      if (waiting > 0 && waiting >= expected) {
        for (int t = 0; t < g_nthreads; t++)
          if (g_fibers[t].state == WAIT_NAMED && g_fibers[t].named_id == id) g_fibers[t].state = RUNNABLE;
        released = true;
      }
    }
    if (!ran && !released) {
      std::fprintf(stderr, "simt_emu: deadlock in block (%u, %u, %u): %d live threads wait at barriers that cannot complete\n",
                   blockIdx.x, blockIdx.y, blockIdx.z, live);
      abort();
    }
  }
}

}  // namespace

extern "C" void emu_flush_stream_for_nccl(void *stream);  // defined with the stream queues below

namespace emu {

void launch(dim3 grid, dim3 block, size_t dynamic_smem_bytes, cudaStream_t stream, const std::function<void()> &thread_body) {
  emu_flush_stream_for_nccl(stream);  // a kernel runs behind everything queued on its stream (MIF_EMU_LAZY_COPIES)
  const size_t n = (size_t)block.x * block.y * block.z;
  if (n == 0 || n > kMaxThreads) {
    std::fprintf(stderr, "simt_emu: block of %zu threads\n", n);
    abort();
  }
  if (dynamic_smem_bytes > 227 * 1024) {  // the opt-in maximum of an sm_100 CTA: a real launch fails with "invalid argument"
    std::fprintf(stderr, "simt_emu: launch with %zu bytes of dynamic shared memory (limit 232448)\n", dynamic_smem_bytes);
    abort();
  }
  if ((size_t)grid.x * grid.y * grid.z == 0) {
    std::fprintf(stderr, "simt_emu: launch with an empty grid (a real launch fails with \"invalid configuration\")\n");
    abort();
  }
  if (grid.y > 65535 || grid.z > 65535) {
    std::fprintf(stderr, "simt_emu: grid (%u, %u, %u) exceeds the y / z limit of 65535\n", grid.x, grid.y, grid.z);
    abort();
  }
  g_nthreads = (int)n;
  g_body = &thread_body;
  blockDim = block;
  gridDim = grid;
  g_dynamic_smem.assign(dynamic_smem_bytes + 1024, 0);
  int t = 0;
  for (unsigned z = 0; z < block.z; z++)
    for (unsigned y = 0; y < block.y; y++)
      for (unsigned x = 0; x < block.x; x++) g_tid[t++] = uint3{x, y, z};
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx = uint3{bx, by, bz};
        // shared memory is uninitialised on the device: poison it with NaNs so stale reads show up
        for (size_t i = 0; i + 8 <= g_dynamic_smem.size(); i += 8) {
          const uint64_t nan_bits = 0x7ff8dead00000000ull;
          memcpy(&g_dynamic_smem[i], &nan_bits, 8);
        }
        run_block();
      }
  g_body = nullptr;
}

void *dynamic_smem() {
  uintptr_t p = reinterpret_cast<uintptr_t>(g_dynamic_smem.data());
  return reinterpret_cast<void *>((p + 1023) & ~uintptr_t(1023));  // like the shared window of a CTA without static shared memory
}

// ---- TMA / mbarrier stand-ins (csrc/mif_tma.cuh) -----------------------------------------------------------------
// Copies are performed as LATE as the programming model allows, so that a kernel that skips a wait or reuses a buffer
// too early computes with NaNs or stale data instead of passing by luck: a load poisons its destination when it is
// issued and is carried out when a thread first tests its barrier; a store is carried out when its issuing thread
// waits for the group (or flagged as an error when the block ends with stores still pending).
namespace {
struct PendingLoad { void *dst; CUtensorMap map; int c[3]; uint64_t *bar; };
struct PendingStore { CUtensorMap map; const void *src; int c[3]; int thread; void *bulk_dst; unsigned bulk_bytes; };
struct BarState { uint64_t *bar; int count, pending, phase; long tx; };
std::vector<PendingLoad> g_loads;
std::vector<PendingStore> g_stores;
std::vector<BarState> g_bars;
long g_spins = 0;
BarState &bar_state(uint64_t *bar) {
  for (BarState &b : g_bars)
    if (b.bar == bar) return b;
  std::fprintf(stderr, "simt_emu: mbarrier %p used before mbarrier.init\n", (void *)bar);
  abort();
}
size_t box_bytes(const CUtensorMap &m) { return (size_t)m.box[0] * m.box[1] * m.box[2] * 8; }
// smem <-> global box copy, dimension 0 fastest in shared memory, rows dense; the swizzle XORs the shared ADDRESS bits
// 4-5 with bits 7-8 (CU_TENSOR_MAP_SWIZZLE_64B as measured on a B200, scripts/probes/tma_probe.cu)
void copy_box(const CUtensorMap &m, char *smem, const int c[3], bool to_smem) {
  for (uint32_t r = 0; r < m.box[1]; r++)
    for (uint32_t x = 0; x < m.box[0]; x++) {
      uintptr_t s = reinterpret_cast<uintptr_t>(smem) + ((size_t)r * m.box[0] + x) * 8;
      s ^= ((s >> 7) & m.swizzle_mask) << 4;
      double *sp = reinterpret_cast<double *>(s);
      const uint64_t g0 = (uint64_t)(c[0] + (long)x), g1 = (uint64_t)(c[1] + (long)r), g2 = (uint64_t)c[2];
      const bool inside = c[0] + (long)x >= 0 && c[1] + (long)r >= 0 && c[2] >= 0 && g0 < m.dim[0] && g1 < m.dim[1] && g2 < m.dim[2];
      double *gp = reinterpret_cast<double *>(static_cast<char *>(m.base) + g0 * m.stride_bytes[0] + g1 * m.stride_bytes[1] + g2 * m.stride_bytes[2]);
      if (to_smem) *sp = inside ? *gp : 0.0;
      else if (inside) *gp = *sp;
    }
}
void check_smem_range(const void *p, size_t bytes, const char *what) {
  const unsigned char *lo = g_dynamic_smem.data(), *hi = lo + g_dynamic_smem.size();
  const unsigned char *q = static_cast<const unsigned char *>(p);
  if (q < lo || q + bytes > hi || (reinterpret_cast<uintptr_t>(p) & 127)) {
    std::fprintf(stderr, "simt_emu: %s: shared-memory box outside the dynamic window or not 128-byte aligned\n", what);
    abort();
  }
}
}  // namespace

void mbar_init(uint64_t *bar, int count) {
  for (size_t i = 0; i < g_bars.size(); i++)
    if (g_bars[i].bar == bar) g_bars.erase(g_bars.begin() + i--);
  g_bars.push_back(BarState{bar, count, count, 0, 0});
}
void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
  BarState &b = bar_state(bar);
  b.tx += bytes;
  b.pending -= 1;
  if (b.pending < 0) {
    std::fprintf(stderr, "simt_emu: more arrivals than the mbarrier's count\n");
    abort();
  }
}
bool mbar_test(uint64_t *bar, unsigned parity) {
  BarState &b = bar_state(bar);
  for (size_t i = 0; i < g_loads.size(); i++)
    if (g_loads[i].bar == bar) {
      copy_box(g_loads[i].map, static_cast<char *>(g_loads[i].dst), g_loads[i].c, true);
      b.tx -= (long)box_bytes(g_loads[i].map);
      g_loads.erase(g_loads.begin() + i--);
    }
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1;
    b.pending = b.count;
  }
  return (unsigned)b.phase != (parity & 1u);
}
void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  check_smem_range(smem_dst, box_bytes(*map), "tma load");
  if (map->swizzle_mask && (reinterpret_cast<uintptr_t>(smem_dst) & 1023)) {
    std::fprintf(stderr, "simt_emu: swizzled box not 1024-byte aligned\n");
    abort();
  }
  bar_state(bar);
  const uint64_t nan_bits = 0x7ff8abcd00000000ull;  // in flight: contents undefined
  for (size_t i = 0; i < box_bytes(*map); i += 8) memcpy(static_cast<char *>(smem_dst) + i, &nan_bits, 8);
  g_loads.push_back(PendingLoad{smem_dst, *map, {c0, c1, c2}, bar});
}
void tma_store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2) {
  check_smem_range(smem_src, box_bytes(*map), "tma store");
  if (map->swizzle_mask && (reinterpret_cast<uintptr_t>(smem_src) & 1023)) {
    std::fprintf(stderr, "simt_emu: swizzled box not 1024-byte aligned\n");
    abort();
  }
  g_stores.push_back(PendingStore{*map, smem_src, {c0, c1, c2}, g_current, nullptr, 0});
}
void tma_store_bulk(void *global_dst, const void *smem_src, unsigned bytes) {
  const unsigned char *lo = g_dynamic_smem.data(), *hi = lo + g_dynamic_smem.size();
  const unsigned char *q = static_cast<const unsigned char *>(smem_src);
  if (q < lo || q + bytes > hi || (reinterpret_cast<uintptr_t>(smem_src) & 15) || (reinterpret_cast<uintptr_t>(global_dst) & 15) || (bytes & 15)) {
    std::fprintf(stderr, "simt_emu: bulk store: misaligned or outside the dynamic shared-memory window\n");
    abort();
  }
  g_stores.push_back(PendingStore{CUtensorMap{}, smem_src, {0, 0, 0}, g_current, global_dst, bytes});
}
void tma_commit_group() {}
void tma_wait_group(int, bool) {
  for (size_t i = 0; i < g_stores.size(); i++)
    if (g_stores[i].thread == g_current) {
      if (g_stores[i].bulk_dst) memcpy(g_stores[i].bulk_dst, g_stores[i].src, g_stores[i].bulk_bytes);
      else copy_box(g_stores[i].map, const_cast<char *>(static_cast<const char *>(g_stores[i].src)), g_stores[i].c, false);
      g_stores.erase(g_stores.begin() + i--);
    }
}
void spin_yield() {
  if (++g_spins > 100000000L) {
    std::fprintf(stderr, "simt_emu: an mbarrier wait never completes\n");
    abort();
  }
  yield(RUNNABLE);
}
void tma_block_end() {
  if (!g_stores.empty() || !g_loads.empty()) {
    std::fprintf(stderr, "simt_emu: block (%u, %u, %u) ended with %zu bulk stores and %zu bulk loads still in flight\n", blockIdx.x,
                 blockIdx.y, blockIdx.z, g_stores.size(), g_loads.size());
    abort();
  }
  g_bars.clear();
  g_spins = 0;
}
void block_barrier() { yield(WAIT_BLOCK); }
void warp_barrier() { yield(WAIT_WARP); }
void named_barrier(int id, int count) {
  g_fibers[g_current].named_id = id;
  g_fibers[g_current].named_count = count;
  yield(WAIT_NAMED);
}
void *shuffle_slot(int lane, int parity) { return &g_shuffle[g_current / 32][parity][lane & 31]; }
int lane_id() { return g_current & 31; }
int &shuffle_parity() { return g_parity[g_current]; }

}  // namespace emu

// ---- CUDA runtime stand-ins: "device" memory is host memory --------------------------------------------------
// Streams.  By default everything is carried out at once, in program order.  MIF_EMU_LAZY_COPIES=1 models the freedom a
// device has with ASYNCHRONOUS COPIES: cudaMemcpyAsync / cudaMemcpy3DAsync / cudaMemsetAsync, event records and
// cudaStreamWaitEvent are queued per stream and carried out as LATE as the programming model allows -- when the host
// synchronises with the stream (or the device), when an event recorded behind them is waited for (by the host, or by
// another stream at the moment THAT stream's wait is carried out), or when a kernel / an NCCL operation is issued on
// the same stream (those run at once, after the stream's own queue).  A missing event dependency then shows: the
// consumer computes with data the copy has not delivered yet, or a late copy reads a buffer that was reused too early.
struct StreamOp {
  std::function<void()> run;
  uint64_t marker;  // != 0: the record of an event
};
struct emu_stream {
  int side = 0;  // 1: created with a priority; fake_nccl.cpp reads this int (must stay the first member)
  int index = -1;  // creation order within the process (the legacy default stream: -1)
  std::deque<StreamOp> pending;
};
struct emu_event {
  std::chrono::steady_clock::time_point when;
  cudaStream_t recorded_on = nullptr;
  uint64_t marker = 0;
};

namespace {
emu_stream g_null_stream;
std::vector<emu_stream *> g_streams;  // live streams (cudaDeviceSynchronize, cudaFree)
uint64_t g_next_marker = 1;
// MIF_EMU_LAZY_COPIES: unset = everything at once; "1" / "all" = every stream as late as possible; "odd" / "even" = the
// streams with an odd / even creation index run at once and the others as late as possible -- one stream racing ahead
// of another is what exposes a buffer that is refilled before its previous contents were consumed.
int lazy_mode() {
  static const int mode = [] {
    const char *v = getenv("MIF_EMU_LAZY_COPIES");
    if (!v) return 0;
    if (!strcmp(v, "odd")) return 2;
    if (!strcmp(v, "even")) return 3;
    return 1;
  }();
  return mode;
}
bool lazy_copies() { return lazy_mode() != 0; }
bool trace_streams() {  // MIF_EMU_TRACE=1: log the queueing and the carrying out of stream operations
  static const bool on = getenv("MIF_EMU_TRACE") != nullptr;
  return on;
}
int g_stream_count = 0;
emu_stream *queue_of(cudaStream_t s) { return s ? s : &g_null_stream; }
// Carry out the queued operations of a stream in order: all of them, or up to and including the record with this marker
// (nothing if that record has been carried out already).
void flush_stream(cudaStream_t stream, uint64_t up_to_marker = 0) {
  emu_stream *q = queue_of(stream);
  if (up_to_marker) {
    bool found = false;
    for (const StreamOp &op : q->pending) found = found || op.marker == up_to_marker;
    if (!found) return;
  }
  while (!q->pending.empty()) {
    StreamOp op = std::move(q->pending.front());
    q->pending.pop_front();
    if (trace_streams()) std::fprintf(stderr, "simt_emu:   carry out an operation of stream %d (event record %llu)\n", q->index, (unsigned long long)op.marker);
    op.run();  // may flush other streams (a queued cudaStreamWaitEvent)
    if (up_to_marker && op.marker == up_to_marker) return;
  }
}
void flush_all_streams() {
  flush_stream(nullptr);
  for (size_t i = 0; i < g_streams.size(); i++) flush_stream(g_streams[i]);
}
void enqueue(cudaStream_t stream, std::function<void()> run, uint64_t marker = 0) {
  emu_stream *q = queue_of(stream);
  if (trace_streams()) std::fprintf(stderr, "simt_emu: queue on stream %d (event record %llu), %zu pending\n", q->index, (unsigned long long)marker, q->pending.size());
  q->pending.push_back(StreamOp{std::move(run), marker});
  const bool eager = (lazy_mode() == 2 && q->index >= 0 && (q->index & 1)) || (lazy_mode() == 3 && q->index >= 0 && !(q->index & 1));
  if (eager) flush_stream(stream);
}
}  // namespace
// for fake_nccl.cpp: an NCCL operation issued on `stream` runs behind everything queued on that stream
extern "C" void emu_flush_stream_for_nccl(void *stream) {
  if (lazy_copies()) flush_stream(static_cast<cudaStream_t>(stream));
}

// fake_nccl.cpp's fake_nccl_flush, if that library is loaded in this process (MIFGPU_NCCL_LIB)
static void flush_late_exchanges(cudaStream_t stream) {
  static int (*flush)(void *) = nullptr;
  static bool looked = false;
  if (!looked || !flush) {
    const char *path = getenv("MIFGPU_NCCL_LIB");
    void *handle = path ? dlopen(path, RTLD_NOW | RTLD_NOLOAD) : nullptr;
    if (handle) flush = reinterpret_cast<int (*)(void *)>(dlsym(handle, "fake_nccl_flush"));
    if (handle)
      if (auto set = reinterpret_cast<void (*)(void (*)(void *))>(dlsym(handle, "fake_nccl_set_stream_flush")))
        set(emu_flush_stream_for_nccl);
    looked = true;
  }
  if (flush) flush(stream);
}

// With MIF_SIMT_IPC=1 (multi-process runs) every allocation is a shared-memory file, so that cudaIpcGetMemHandle /
// cudaIpcOpenMemHandle can map a buffer of another rank's process: the handle carries the file name.
namespace {
struct Allocation { void *ptr; size_t bytes; std::string shm_name; bool mapped_peer; };
std::vector<Allocation> g_allocations;
bool ipc_mode() {
  static const bool on = getenv("MIF_SIMT_IPC") != nullptr;
  return on;
}
void unlink_all() {
  for (Allocation &a : g_allocations)
    if (!a.shm_name.empty() && !a.mapped_peer) shm_unlink(a.shm_name.c_str());
}
}  // namespace

cudaError_t emu_malloc(void **ptr, size_t bytes) {
  const size_t rounded = bytes ? (bytes + 4095) / 4096 * 4096 : 4096;
  std::string name;
  if (ipc_mode()) {
    static int counter = 0;
    static bool registered = false;
    if (!registered) {
      atexit(unlink_all);
      registered = true;
    }
    char buf[64];
    std::snprintf(buf, sizeof(buf), "/mif_simt_%d_%d", (int)getpid(), counter++);
    name = buf;
    const int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)rounded) != 0) return cudaErrorEmu;
    *ptr = mmap(nullptr, rounded, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (*ptr == MAP_FAILED) return cudaErrorEmu;
  } else {
    *ptr = aligned_alloc(4096, rounded);
    if (!*ptr) return cudaErrorEmu;
  }
  // device memory is uninitialised: NaN pattern
  const uint64_t nan_bits = 0x7ff8beef00000000ull;
  for (size_t i = 0; i + 8 <= rounded; i += 8) memcpy(static_cast<char *>(*ptr) + i, &nan_bits, 8);
  g_allocations.push_back(Allocation{*ptr, rounded, name, false});
  return cudaSuccess;
}
cudaError_t cudaFree(void *ptr) {
  if (!ptr) return cudaSuccess;
  flush_all_streams();  // cudaFree synchronises the device
  for (size_t i = 0; i < g_allocations.size(); i++)
    if (g_allocations[i].ptr == ptr) {
      if (g_allocations[i].shm_name.empty()) free(ptr);
      else {
        munmap(ptr, g_allocations[i].bytes);
        shm_unlink(g_allocations[i].shm_name.c_str());
      }
      g_allocations.erase(g_allocations.begin() + i);
      return cudaSuccess;
    }
  return cudaErrorEmu;
}
cudaError_t cudaFreeHost(void *ptr) { return cudaFree(ptr); }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t stream) {
  if (lazy_copies()) enqueue(stream, [=] { memmove(dst, src, bytes); });
  else memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemset(void *dst, int value, size_t bytes) { memset(dst, value, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t stream) {
  if (lazy_copies()) enqueue(stream, [=] { memset(dst, value, bytes); });
  else memset(dst, value, bytes);
  return cudaSuccess;
}
static void copy_3d(const cudaMemcpy3DParms *p) {
  const cudaPitchedPtr &s = p->srcPtr, &d = p->dstPtr;
  for (size_t z = 0; z < p->extent.depth; z++)
    for (size_t y = 0; y < p->extent.height; y++) {
      const char *src = static_cast<const char *>(s.ptr) + ((p->srcPos.z + z) * s.ysize + p->srcPos.y + y) * s.pitch + p->srcPos.x;
      char *dst = static_cast<char *>(d.ptr) + ((p->dstPos.z + z) * d.ysize + p->dstPos.y + y) * d.pitch + p->dstPos.x;
      memmove(dst, src, p->extent.width);
    }
}
cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms *p, cudaStream_t stream) {
  if (lazy_copies()) {
    const cudaMemcpy3DParms parms = *p;
    enqueue(stream, [parms] { copy_3d(&parms); });
  } else {
    copy_3d(p);
  }
  return cudaSuccess;
}
cudaError_t cudaGetDeviceCount(int *count) { *count = 64; return cudaSuccess; }  // any ordinal a rank asks for exists
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *device) { *device = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *value, int attr, int) {
  *value = (attr == cudaDevAttrMultiProcessorCount) ? 3 : 0;  // few "SMs": persistent kernels loop over several tiles per CTA
  return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() {
  flush_all_streams();
  flush_late_exchanges(nullptr);
  return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t err) { return err == cudaSuccess ? "no error" : "simt_emu: unsupported call"; }
cudaError_t cudaStreamCreate(cudaStream_t *stream) {
  *stream = new emu_stream();
  (*stream)->index = g_stream_count++;
  g_streams.push_back(*stream);
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *stream, unsigned) { return cudaStreamCreate(stream); }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *stream, unsigned, int) { cudaStreamCreate(stream); (*stream)->side = 1; return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) { if (least) *least = 0; if (greatest) *greatest = -5; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t stream) {
  flush_stream(stream);  // work already issued completes
  for (size_t i = 0; i < g_streams.size(); i++)
    if (g_streams[i] == stream) g_streams.erase(g_streams.begin() + i);
  delete stream;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) {
  flush_stream(s);
  flush_late_exchanges(s);
  return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *event) { *event = new emu_event(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *event, unsigned) { *event = new emu_event(); return cudaSuccess; }
// Kernels and copies are synchronous here; only the stand-in NCCL can defer the exchanges of a side stream
// (MIF_FAKE_NCCL_LATE, fake_nccl.cpp) -- they complete when another stream waits for an event recorded behind them.
cudaError_t cudaStreamWaitEvent(cudaStream_t stream, cudaEvent_t event, unsigned) {
  if (!event || !event->marker) return cudaSuccess;  // never recorded: nothing to wait for
  const cudaStream_t source = event->recorded_on;
  const uint64_t marker = event->marker;  // the record the wait refers to is the one made so far
  auto wait = [source, marker] {
    flush_stream(source, marker);
    if (source) flush_late_exchanges(source);
  };
  if (lazy_copies()) enqueue(stream, wait);
  else wait();
  return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t event) { delete event; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t event, cudaStream_t s) {
  event->recorded_on = s;
  event->marker = g_next_marker++;
  if (lazy_copies()) enqueue(s, [event] { event->when = std::chrono::steady_clock::now(); }, event->marker);
  else event->when = std::chrono::steady_clock::now();
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t event) {
  if (event && event->marker) flush_stream(event->recorded_on, event->marker);
  return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t start, cudaEvent_t stop) {
  cudaEventSynchronize(start);
  cudaEventSynchronize(stop);
  *ms = std::chrono::duration<float, std::milli>(stop->when - start->when).count();
  return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *handle, void *ptr) {
  for (const Allocation &a : g_allocations)
    if (a.ptr == ptr && !a.shm_name.empty()) {
      memset(handle, 0, sizeof(*handle));
      std::snprintf(handle->reserved, sizeof(handle->reserved), "%s", a.shm_name.c_str());
      return cudaSuccess;
    }
  return cudaErrorEmu;
}
cudaError_t cudaIpcOpenMemHandle(void **ptr, cudaIpcMemHandle_t handle, unsigned) {
  handle.reserved[sizeof(handle.reserved) - 1] = 0;
  const int fd = shm_open(handle.reserved, O_RDWR, 0600);
  if (fd < 0) return cudaErrorEmu;
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    return cudaErrorEmu;
  }
  *ptr = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (*ptr == MAP_FAILED) return cudaErrorEmu;
  g_allocations.push_back(Allocation{*ptr, (size_t)st.st_size, handle.reserved, true});
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *ptr) {
  for (size_t i = 0; i < g_allocations.size(); i++)
    if (g_allocations[i].ptr == ptr && g_allocations[i].mapped_peer) {
      munmap(ptr, g_allocations[i].bytes);
      g_allocations.erase(g_allocations.begin() + i);
      return cudaSuccess;
    }
  return cudaErrorEmu;
}
