// mif_api.cu -- the C ABI of libmifgpu (include/mifgpu.h): context, tensors and the orchestration of one
// projection time step (src/Timestep.cpp:97-156) on one CUDA stream.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mifgpu.h"
#include "mif_kernels.h"

using namespace mifgpu;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// NCCL is bound at run time (dlopen) the first time a multi-GPU context is requested: single-GPU users need no NCCL
// at all, and inside a process that already carries an NCCL (e.g. PyTorch's bundled one) that copy is reused
// instead of a second libnccl.so.2 being forced into the process at load time.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
constexpr ncclDataType_t kNcclReal = sizeof(real) == 8 ? ncclDouble : ncclFloat;  // MPI_MIF_REAL (include/Real.h:12,16)

bool load_nccl() {
  if (g_nccl.ok) return true;
  // MIFGPU_NCCL_LIB names the library to bind instead (a system NCCL next to a framework's bundled one; the tests'
  // in-process stand-in on machines without GPUs).
  const char *override_path = getenv("MIFGPU_NCCL_LIB");
  void *handle = override_path ? dlopen(override_path, RTLD_NOW | RTLD_LOCAL) : dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!handle && !override_path) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!handle) {
    g_last_error = std::string("cannot load ") + (override_path ? override_path : "libnccl.so.2") + ": " + dlerror();
    return false;
  }
  auto sym = [&](const char *name) { return dlsym(handle, name); };
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
  g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(sym("ncclGroupStart"));
  g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(sym("ncclGroupEnd"));
  g_nccl.Send = reinterpret_cast<decltype(g_nccl.Send)>(sym("ncclSend"));
  g_nccl.Recv = reinterpret_cast<decltype(g_nccl.Recv)>(sym("ncclRecv"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.GroupStart && g_nccl.GroupEnd &&
              g_nccl.Send && g_nccl.Recv && g_nccl.AllReduce && g_nccl.GetErrorString;
  if (!g_nccl.ok) g_last_error = "libnccl.so.2 lacks a required symbol";
  return g_nccl.ok;
}

#define NCCL_TRY(expr)                                                                            \
  do {                                                                                            \
    ncclResult_t err__ = (expr);                                                                  \
    if (err__ != ncclSuccess)                                                                     \
      return fail(MIFGPU_ERR_COMM, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

#define CUDA_TRY(expr)                                                                            \
  do {                                                                                            \
    cudaError_t err__ = (expr);                                                                   \
    if (err__ != cudaSuccess)                                                                     \
      return fail(MIFGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

}  // namespace

enum ProfCategory {
  PROF_STAGE1, PROF_STAGE2, PROF_STAGE3, PROF_BC, PROF_DIVERGENCE, PROF_NHN, PROF_SWEEP_X_FWD, PROF_SWEEP_Y_FWD,
  PROF_SWEEP_Z, PROF_SWEEP_Y_INV, PROF_SWEEP_X_INV, PROF_PERIODIC, PROF_CORRECT, PROF_HALO, PROF_TRANSPOSE, PROF_COUNT
};
const char *const kProfNames[PROF_COUNT] = {"stage1", "stage2", "stage3", "bc_faces", "divergence", "nhn_rhs",
                                            "sweep_x_fwd", "sweep_y_fwd", "sweep_z_fused", "sweep_y_inv",
                                            "sweep_x_inv", "periodic", "correct", "halo_exchange",
                                            "transpose_alltoall"};
struct ProfRecord {
  int category;
  cudaEvent_t start, stop;
};

struct mifgpu_ctx {
  mifgpu_params params;
  Geom g;
  int n_points[3];
  cudaStream_t stream = nullptr;
  PoissonPlan *plan = nullptr;
  uint64_t launches = 0;
  bool profiling = false;
  std::vector<ProfRecord> prof_records;
  // multi-GPU slab decomposition (Py = 1, Pz = nranks): NCCL communicator, y ranges of the z pencils, buffers
  ncclComm_t comm = nullptr;
  int nranks = 1;
  std::vector<int> ylo;      // nranks + 1: y rows [ylo[r], ylo[r+1]) of the transform domain belong to rank r's z pencil
  std::vector<int> zlo;      // nranks + 1: owner z points [zlo[r], zlo[r+1]) of rank r's slab
  int *ylo_dev = nullptr;
  real *xfer = nullptr;    // send / receive staging, one local owner volume
  real *zbuf = nullptr;    // z pencil: zbuf[z][y_local][x]
  // peer-memory transposes: zbuf / xfer of every rank mapped with CUDA IPC (index = rank; own entry = local pointer)
  bool peer_mode = false;
  real *zbuf_peer[8] = {};
  real *xfer_peer[8] = {};
  int *barrier_word = nullptr;
  // Halo exchanges overlapped with the interior planes of the kernel that consumes them (z slabs): the NCCL plane
  // exchange runs on comm_stream after everything queued on the compute stream so far (ev_fork); the consumer kernel is
  // launched for its interior planes, the compute stream then waits for ev_join and the boundary planes follow.
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap_halos = false;
  bool halo_pending = false;
  // MIFGPU_REFERENCE_HALOS=1: the y / z exchanges of a Py x Pz run exactly as the reference orders them, stale edge
  // ghosts included (exchange_halos)
  bool reference_halos = false;
  // Py > 1: pencil decomposition in the 2Decomp layout (deps/2Decomp_C/C2Decomp.cpp:241-423): x pencil (this rank's
  // sub-domain) -> y pencil (x distributed over the Py ranks of the same z_rank) -> z pencil (y distributed over the
  // Pz ranks of the same y_rank).  Block distributions as in src/Constants.cpp:78-79 (bigger blocks on the low ranks).
  int Py = 1, Pz = 1, y_rank = 0, z_rank = 0;
  std::vector<int> ys, xs;   // Py + 1: owner y rows of y_rank r; x columns of the y / z pencils of y_rank r
  std::vector<int> zs, yzs;  // Pz + 1: owner z planes of z_rank r; y rows of the z pencil of z_rank r
  real *ypen = nullptr, *zpen = nullptr;           // pencil buffers, rows of pen_pitch doubles
  real *box_send = nullptr, *box_recv = nullptr;   // compact staging of the box exchanges
  size_t box_capacity = 0;                           // doubles in each of them
  int pen_pitch = 0;
  real *staging = nullptr; // compact device copy of one tensor for host transfers
  size_t staging_bytes = 0;
  // asynchronous host transfers (mifgpu_tensor_upload_async / _download_async): one stream and one compact staging
  // buffer per direction, so that H2D, D2H and the kernels of independent tensors overlap
  // Asynchronous transfers: [0] host -> device and [1] device -> host over PCIe, [2] / [3] the re-pitching copies on the
  // device behind / ahead of them.  Two compact staging buffers per direction, so that the link moves tensor n + 1 while
  // tensor n is re-pitched; ev_stage_filled / ev_stage_free hand a buffer from one stream to the other and back.
  cudaStream_t copy_stream[4] = {nullptr, nullptr, nullptr, nullptr};
  real *copy_staging[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  size_t copy_staging_bytes[2] = {0, 0};
  cudaEvent_t ev_stage_filled[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  cudaEvent_t ev_stage_free[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  int copy_next[2] = {0, 0};
  // host-callback boundary faces: pinned staging + device copies, [which][component][face]
  real *face_host[2][3][6] = {};
  real *face_dev[2][3][6] = {};
};

struct mifgpu_tensor {
  mifgpu_ctx *ctx;
  int staggering;
  real *data;
  // ordering between the three streams: last asynchronous upload into / download out of this tensor, last compute
  // call that used it (waiting on an event that was never recorded is a no-op)
  cudaEvent_t ev_uploaded = nullptr, ev_downloaded = nullptr, ev_computed = nullptr;
};

namespace {

// Derived constants, same formulas as src/Constants.cpp:63-101.
int build_geometry(const mifgpu_params &p, Geom &g, int n_points[3]) {
  if (p.Nx_global < 2 || p.Ny_global < 2 || p.Nz_global < 2) return fail(MIFGPU_ERR_INVALID, "grid needs >= 2 points per direction");
  if (p.num_time_steps == 0 || !(p.final_time > 0)) return fail(MIFGPU_ERR_INVALID, "num_time_steps and final_time must be positive");
  if (p.Py < 1 || p.Pz < 1) return fail(MIFGPU_ERR_INVALID, "Py and Pz must be >= 1");
  if (p.rank < 0 || p.rank >= p.Py * p.Pz) return fail(MIFGPU_ERR_INVALID, "rank outside [0, Py*Pz)");
  const bool per[3] = {p.periodic_bc[0] != 0, p.periodic_bc[1] != 0, p.periodic_bc[2] != 0};
  const int Py = p.Py, Pz = p.Pz, rank = p.rank;
  const int y_rank = rank / Pz, z_rank = rank % Pz;
  const size_t ny_pts = p.Ny_global - per[1], nz_pts = p.Nz_global - per[2];
  const size_t Ny_owner = ny_pts / Py + ((size_t)y_rank < ny_pts % Py ? 1 : 0);
  const size_t Nz_owner = nz_pts / Pz + ((size_t)z_rank < nz_pts % Pz ? 1 : 0);
  const size_t Nx = per[0] ? p.Nx_global + 1 : p.Nx_global;
  const size_t Ny = (Py == 1) ? (per[1] ? p.Ny_global + 1 : p.Ny_global)
                              : (((y_rank == Py - 1 && !per[1]) || (y_rank == 0 && !per[1])) ? Ny_owner + 1 : Ny_owner + 2);
  const size_t Nz = (Pz == 1) ? (per[2] ? p.Nz_global + 1 : p.Nz_global)
                              : (((z_rank == Pz - 1 && !per[2]) || (z_rank == 0 && !per[2])) ? Nz_owner + 1 : Nz_owner + 2);
  const size_t Nx_st = per[0] ? Nx : Nx + 1;
  const size_t Ny_st = (y_rank == Py - 1) ? Ny + 1 : Ny;
  const size_t Nz_st = (z_rank == Pz - 1) ? Nz + 1 : Nz;
  if (Nx_st > (1u << 30) || Ny_st > (1u << 30) || Nz_st > (1u << 30)) return fail(MIFGPU_ERR_INVALID, "extent too large");

  g.Nx = (int)Nx; g.Ny = (int)Ny; g.Nz = (int)Nz;
  g.sx[0] = (int)Nx_st; g.sy[0] = (int)Ny;    g.sz[0] = (int)Nz;
  g.sx[1] = (int)Nx;    g.sy[1] = (int)Ny_st; g.sz[1] = (int)Nz;
  g.sx[2] = (int)Nx;    g.sy[2] = (int)Ny;    g.sz[2] = (int)Nz_st;
  g.sx[3] = (int)Nx;    g.sy[3] = (int)Ny;    g.sz[3] = (int)Nz;
  g.PX = (int)((Nx_st + 15) / 16 * 16);
  g.PY = (int)Ny_st;
  g.PZ = (int)Nz_st;
  g.plane = (long long)g.PX * g.PY;
  g.volume = g.plane * g.PZ;
  for (int d = 0; d < 3; d++) g.periodic[d] = per[d];
  g.base_i = per[0] ? -1 : 0;
  g.base_j = (int)(ny_pts / Py * y_rank + std::min((size_t)y_rank, ny_pts % Py)) - ((y_rank > 0 || per[1]) ? 1 : 0);
  g.base_k = (int)(nz_pts / Pz * z_rank + std::min((size_t)z_rank, nz_pts % Pz)) - ((z_rank > 0 || per[2]) ? 1 : 0);
  g.prev_y = (y_rank == 0) ? ((Py > 1 && per[1]) ? rank + (Py - 1) * Pz : -1) : rank - Pz;
  g.next_y = (y_rank == Py - 1) ? ((Py > 1 && per[1]) ? rank - (Py - 1) * Pz : -1) : rank + Pz;
  g.prev_z = (z_rank == 0) ? ((Pz > 1 && per[2]) ? rank + Pz - 1 : -1) : rank - 1;
  g.next_z = (z_rank == Pz - 1) ? ((Pz > 1 && per[2]) ? rank - (Pz - 1) : -1) : rank + 1;
  // owner ranges of pressure points (include/StaggeredTensorMacros.h:41-83)
  g.own_lo[0] = per[0] ? 1 : 0;
  g.own_hi[0] = per[0] ? g.Nx - 1 : g.Nx;
  g.own_lo[1] = (g.prev_y != -1 || per[1]) ? 1 : 0;
  g.own_hi[1] = (g.next_y != -1 || per[1]) ? g.Ny - 1 : g.Ny;
  g.own_lo[2] = (g.prev_z != -1 || per[2]) ? 1 : 0;
  g.own_hi[2] = (g.next_z != -1 || per[2]) ? g.Nz - 1 : g.Nz;

  g.min_x = p.min_x_global; g.min_y = p.min_y_global; g.min_z = p.min_z_global;
  const size_t Nx_domains = p.Nx_global - 1, Ny_domains = p.Ny_global - 1, Nz_domains = p.Nz_global - 1;
  g.dt = p.final_time / p.num_time_steps;
  g.dx = p.x_size / Nx_domains; g.dy = p.y_size_global / Ny_domains; g.dz = p.z_size_global / Nz_domains;
  g.one_over_2_dx = 1 / (2 * g.dx); g.one_over_2_dy = 1 / (2 * g.dy); g.one_over_2_dz = 1 / (2 * g.dz);
  g.one_over_8_dx = 1 / (8 * g.dx); g.one_over_8_dy = 1 / (8 * g.dy); g.one_over_8_dz = 1 / (8 * g.dz);
  g.one_over_dx2_Re = 1 / (p.Re * g.dx * g.dx);
  g.one_over_dy2_Re = 1 / (p.Re * g.dy * g.dy);
  g.one_over_dz2_Re = 1 / (p.Re * g.dz * g.dz);
  g.dx_over_2 = g.dx / 2; g.dy_over_2 = g.dy / 2; g.dz_over_2 = g.dz / 2;
  g.one_over_dx = 1 / g.dx; g.one_over_dy = 1 / g.dy; g.one_over_dz = 1 / g.dz;
  n_points[0] = (int)(p.Nx_global - per[0]);
  n_points[1] = (int)ny_pts;
  n_points[2] = (int)nz_pts;
  return MIFGPU_OK;
}

// Optional per-kernel timing with CUDA events on the context's stream (mifgpu_profile_*).
struct ProfScope {
  mifgpu_ctx *ctx;
  ProfRecord rec;
  bool active;
  cudaStream_t stream;
  ProfScope(mifgpu_ctx *c, int category, cudaStream_t on = nullptr) : ctx(c), active(c->profiling), stream(on ? on : c->stream) {
    if (!active) return;
    rec.category = category;
    cudaEventCreate(&rec.start);
    cudaEventCreate(&rec.stop);
    cudaEventRecord(rec.start, stream);
  }
  ~ProfScope() {
    if (!active) return;
    cudaEventRecord(rec.stop, stream);
    ctx->prof_records.push_back(rec);
  }
};

Vec3 vec3(mifgpu_tensor *const t[3]) { return Vec3{{t[0]->data, t[1]->data, t[2]->data}}; }
CVec3 cvec3(mifgpu_tensor *const t[3]) { return CVec3{{t[0]->data, t[1]->data, t[2]->data}}; }

int check_triple(const mifgpu_ctx *ctx, mifgpu_tensor *const t[3], const char *name) {
  if (!t) return fail(MIFGPU_ERR_INVALID, "%s is NULL", name);
  for (int c = 0; c < 3; c++) {
    if (!t[c]) return fail(MIFGPU_ERR_INVALID, "%s[%d] is NULL", name, c);
    if (t[c]->ctx != ctx) return fail(MIFGPU_ERR_INVALID, "%s[%d] belongs to another context", name, c);
    if (t[c]->staggering != c) return fail(MIFGPU_ERR_INVALID, "%s[%d] has staggering %d", name, c, t[c]->staggering);
  }
  return MIFGPU_OK;
}

int check_scalar(const mifgpu_ctx *ctx, const mifgpu_tensor *t, const char *name) {
  if (!t) return fail(MIFGPU_ERR_INVALID, "%s is NULL", name);
  if (t->ctx != ctx) return fail(MIFGPU_ERR_INVALID, "%s belongs to another context", name);
  if (t->staggering != MIFGPU_STAGGER_NONE) return fail(MIFGPU_ERR_INVALID, "%s must be unstaggered", name);
  return MIFGPU_OK;
}

// Compute entry points bracket their work with this: the context's stream first waits for asynchronous host transfers
// of the tensors it is about to use, and on the way out every tensor remembers the point of the stream after which
// it may be overwritten (upload) or read (download) by the copy streams.
struct UseScope {
  mifgpu_ctx *ctx;
  std::vector<mifgpu_tensor *> tensors;
  explicit UseScope(mifgpu_ctx *c) : ctx(c) {}
  void add(mifgpu_tensor *t) {
    if (!t || !t->ev_uploaded) return;
    cudaStreamWaitEvent(ctx->stream, t->ev_uploaded, 0);
    cudaStreamWaitEvent(ctx->stream, t->ev_downloaded, 0);
    tensors.push_back(t);
  }
  void add(mifgpu_tensor *const t[3]) {
    for (int c = 0; c < 3; c++) add(t[c]);
  }
  ~UseScope() {
    for (mifgpu_tensor *t : tensors) cudaEventRecord(t->ev_computed, ctx->stream);
  }
};

bool face_is_active(const Geom &g, int face) {
  switch (face) {
    case 0: return g.prev_z == -1 && !g.periodic[2];
    case 1: return g.next_z == -1 && !g.periodic[2];
    case 2: return g.prev_y == -1 && !g.periodic[1];
    case 3: return g.next_y == -1 && !g.periodic[1];
    default: return !g.periodic[0];
  }
}

// Fill the device face tables through the host callback.  which = 0: velocity faces of the three
// components; which = 1: Neumann data on the faces of the pressure tensor (component = normal direction).
int fill_face_tables(mifgpu_ctx *ctx, const mifgpu_bc *bc, int which, double time, double time_prev, BcDev &dev) {
  if (!bc->callback) return fail(MIFGPU_ERR_INVALID, "boundary data needs a host callback but bc->callback is NULL");
  const Geom &g = ctx->g;
  for (int comp = 0; comp < 3; comp++) {
    for (int face = 0; face < 6; face++) {
      dev.tables[comp][face] = nullptr;
      const int dir = 2 - face / 2;
      if (which == 1 && comp != dir) continue;
      if (which == 0 && !face_is_active(g, face)) continue;
      const int t = (which == 1) ? 3 : comp;
      const size_t na = (dir == 0) ? g.sy[t] : g.sx[t], nb = (dir == 2) ? g.sy[t] : g.sz[t];
      const size_t bytes = na * nb * sizeof(real);
      if (!ctx->face_host[which][comp][face]) {
        CUDA_TRY(cudaMallocHost(&ctx->face_host[which][comp][face], bytes));
        CUDA_TRY(cudaMalloc(&ctx->face_dev[which][comp][face], bytes));
      }
      // The previous asynchronous copy out of this staging buffer must have finished before it is refilled.
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      bc->callback(bc->user, which, time, time_prev, comp, face, ctx->face_host[which][comp][face]);
      CUDA_TRY(cudaMemcpyAsync(ctx->face_dev[which][comp][face], ctx->face_host[which][comp][face], bytes,
                               cudaMemcpyHostToDevice, ctx->stream));
      dev.tables[comp][face] = ctx->face_dev[which][comp][face];
    }
  }
  return MIFGPU_OK;
}

// 1-cell halo exchange in z of whole (padded) planes: plane 1 -> prev rank's last plane, plane sz-2 -> next rank's
// plane 0 (src/StaggeredTensor.cpp:60-135; tags and Isend/Recv become one NCCL group).
int exchange_z(mifgpu_ctx *ctx, mifgpu_tensor *const *tensors, int count, cudaStream_t on) {
  if (ctx->nranks == 1) return MIFGPU_OK;
  const Geom &g = ctx->g;
  cudaStream_t stream = on ? on : ctx->stream;
  ProfScope prof(ctx, PROF_HALO, stream);
  const size_t plane = (size_t)g.plane;
  NCCL_TRY(g_nccl.GroupStart());
  for (int t = 0; t < count; t++) {
    real *data = tensors[t]->data;
    const int sz = g.sz[tensors[t]->staggering];
    if (g.prev_z != -1 && g.prev_z == g.next_z) {
      // Periodic z on two ranks: both neighbours are the same peer.  NCCL pairs the operations between two ranks in
      // issue order (the reference tells them apart by tag, src/StaggeredTensor.cpp:60-135): the peer's first send is
      // its plane 1, which is this rank's TOP ghost, its second send (plane sz-2) the bottom ghost.
      NCCL_TRY(g_nccl.Send(data + plane, plane, kNcclReal, g.prev_z, ctx->comm, stream));
      NCCL_TRY(g_nccl.Send(data + plane * (sz - 2), plane, kNcclReal, g.next_z, ctx->comm, stream));
      NCCL_TRY(g_nccl.Recv(data + plane * (sz - 1), plane, kNcclReal, g.next_z, ctx->comm, stream));
      NCCL_TRY(g_nccl.Recv(data, plane, kNcclReal, g.prev_z, ctx->comm, stream));
      continue;
    }
    if (g.prev_z != -1) {
      NCCL_TRY(g_nccl.Send(data + plane, plane, kNcclReal, g.prev_z, ctx->comm, stream));
      NCCL_TRY(g_nccl.Recv(data, plane, kNcclReal, g.prev_z, ctx->comm, stream));
    }
    if (g.next_z != -1) {
      NCCL_TRY(g_nccl.Send(data + plane * (sz - 2), plane, kNcclReal, g.next_z, ctx->comm, stream));
      NCCL_TRY(g_nccl.Recv(data + plane * (sz - 1), plane, kNcclReal, g.next_z, ctx->comm, stream));
    }
  }
  NCCL_TRY(g_nccl.GroupEnd());
  return MIFGPU_OK;
}

// Stream-ordered barrier across all ranks (a one-word NCCL all-reduce): every rank's preceding kernels -- including
// their stores into peer memory -- have completed when it returns.
int rank_barrier(mifgpu_ctx *ctx) {
  NCCL_TRY(g_nccl.AllReduce(ctx->barrier_word, ctx->barrier_word + 1, 1, ncclInt, ncclSum, ctx->comm, ctx->stream));
  return MIFGPU_OK;
}

// Map the pencil (zbuf) and slab staging (xfer) buffers of all ranks into this process: the IPC handles travel
// through device buffers with one grouped NCCL exchange.
int setup_peer_memory(mifgpu_ctx *ctx) {
  const int P = ctx->nranks, me = ctx->params.rank;
  struct Handles { cudaIpcMemHandle_t zbuf, xfer; };
  std::vector<Handles> all(P);
  CUDA_TRY(cudaIpcGetMemHandle(&all[me].zbuf, ctx->zbuf));
  CUDA_TRY(cudaIpcGetMemHandle(&all[me].xfer, ctx->xfer));
  Handles *dev = nullptr;
  CUDA_TRY(cudaMalloc(&dev, sizeof(Handles) * P));
  CUDA_TRY(cudaMemcpy(dev + me, &all[me], sizeof(Handles), cudaMemcpyHostToDevice));
  NCCL_TRY(g_nccl.GroupStart());
  for (int r = 0; r < P; r++) {
    if (r == me) continue;
    NCCL_TRY(g_nccl.Send(dev + me, sizeof(Handles), ncclChar, r, ctx->comm, ctx->stream));
    NCCL_TRY(g_nccl.Recv(dev + r, sizeof(Handles), ncclChar, r, ctx->comm, ctx->stream));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(cudaMemcpy(all.data(), dev, sizeof(Handles) * P, cudaMemcpyDeviceToHost));
  cudaFree(dev);
  bool ok = true;
  for (int r = 0; r < P; r++) {
    if (r == me) {
      ctx->zbuf_peer[r] = ctx->zbuf;
      ctx->xfer_peer[r] = ctx->xfer;
      continue;
    }
    void *pz = nullptr, *px = nullptr;
    if (cudaIpcOpenMemHandle(&pz, all[r].zbuf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&px, all[r].xfer, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = false;
      break;
    }
    ctx->zbuf_peer[r] = static_cast<real *>(pz);
    ctx->xfer_peer[r] = static_cast<real *>(px);
  }
  // All ranks must agree, otherwise one side would wait at a barrier the other never reaches.
  int flags[2] = {ok ? 0 : 1, 0};
  CUDA_TRY(cudaMemcpy(ctx->barrier_word, flags, sizeof(flags), cudaMemcpyHostToDevice));
  int rc = rank_barrier(ctx);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(cudaMemcpy(flags, ctx->barrier_word, sizeof(flags), cudaMemcpyDeviceToHost));
  ctx->peer_mode = (flags[1] == 0);
  flags[0] = 0;
  CUDA_TRY(cudaMemcpy(ctx->barrier_word, flags, sizeof(int), cudaMemcpyHostToDevice));
  return MIFGPU_OK;
}

PeerLayout peer_layout(const mifgpu_ctx *ctx) {
  PeerLayout p;
  p.rank = ctx->params.rank;
  p.nranks = ctx->nranks;
  for (int r = 0; r <= ctx->nranks; r++) {
    p.ylo[r] = ctx->ylo[r];
    p.zlo[r] = ctx->zlo[r];
  }
  for (int r = 0; r < ctx->nranks; r++) {
    p.zbuf[r] = ctx->zbuf_peer[r];
    p.xfer[r] = ctx->xfer_peer[r];
  }
  return p;
}

// Slab -> z pencil (forward = true) and back: the 2Decomp Y<->Z transposes (deps/2Decomp_C/TransposeY2Z.cpp:22-50,
// TransposeZ2Y.cpp:18-46) as one grouped NCCL send/recv all-to-all.  Forward: pack rows per destination, receive
// straight into zbuf (the planes of one source are contiguous there).  Backward: send straight out of zbuf,
// receive into the staging buffer, unpack.
int transpose_slab(mifgpu_ctx *ctx, real *field, bool forward) {
  const Geom &g = ctx->g;
  ProfScope prof(ctx, PROF_TRANSPOSE);
  const int me = ctx->params.rank, P = ctx->nranks;
  const long long nz_me = ctx->zlo[me + 1] - ctx->zlo[me], ny_me = ctx->ylo[me + 1] - ctx->ylo[me];
  if (forward) launch_pack_slab(ctx->stream, g, field, ctx->xfer, ctx->ylo_dev, P, &ctx->launches);
  NCCL_TRY(g_nccl.GroupStart());
  for (int r = 0; r < P; r++) {
    const long long ny_r = ctx->ylo[r + 1] - ctx->ylo[r], nz_r = ctx->zlo[r + 1] - ctx->zlo[r];
    real *slab_block = ctx->xfer + nz_me * g.PX * ctx->ylo[r];            // [z_local][y in r's range][x]
    real *pencil_block = ctx->zbuf + (long long)ctx->zlo[r] * ny_me * g.PX;  // planes z in r's slab
    const size_t slab_count = (size_t)(nz_me * ny_r * g.PX), pencil_count = (size_t)(nz_r * ny_me * g.PX);
    if (forward) {
      NCCL_TRY(g_nccl.Send(slab_block, slab_count, kNcclReal, r, ctx->comm, ctx->stream));
      NCCL_TRY(g_nccl.Recv(pencil_block, pencil_count, kNcclReal, r, ctx->comm, ctx->stream));
    } else {
      NCCL_TRY(g_nccl.Send(pencil_block, pencil_count, kNcclReal, r, ctx->comm, ctx->stream));
      NCCL_TRY(g_nccl.Recv(slab_block, slab_count, kNcclReal, r, ctx->comm, ctx->stream));
    }
  }
  NCCL_TRY(g_nccl.GroupEnd());
  if (!forward) launch_unpack_slab(ctx->stream, g, field, ctx->xfer, ctx->ylo_dev, P, &ctx->launches);
  return MIFGPU_OK;
}

// ---- Py > 1: box exchanges ---------------------------------------------------------------------------------------
// A sub-box of a pitched 3-D array (x fastest): element (x, y, z) at base[x + pitch * (y + ysize * z)].
struct Box {
  real *base;
  size_t pitch, ysize;
  int x0, y0, z0, nx, ny, nz;
  size_t count() const { return (size_t)nx * ny * nz; }
};

int copy_box(mifgpu_ctx *ctx, const Box &dst, const Box &src) {
  if (src.count() == 0) return MIFGPU_OK;
  cudaMemcpy3DParms parms;
  std::memset(&parms, 0, sizeof(parms));
  parms.srcPtr = make_cudaPitchedPtr(src.base, src.pitch * sizeof(real), src.pitch, src.ysize);
  parms.srcPos = make_cudaPos((size_t)src.x0 * sizeof(real), (size_t)src.y0, (size_t)src.z0);
  parms.dstPtr = make_cudaPitchedPtr(dst.base, dst.pitch * sizeof(real), dst.pitch, dst.ysize);
  parms.dstPos = make_cudaPos((size_t)dst.x0 * sizeof(real), (size_t)dst.y0, (size_t)dst.z0);
  parms.extent = make_cudaExtent((size_t)src.nx * sizeof(real), (size_t)src.ny, (size_t)src.nz);
  parms.kind = cudaMemcpyDeviceToDevice;
  CUDA_TRY(cudaMemcpy3DAsync(&parms, ctx->stream));
  return MIFGPU_OK;
}

Box compact_box(real *base, const Box &like) {
  return Box{base, (size_t)like.nx, (size_t)like.ny, 0, 0, 0, like.nx, like.ny, like.nz};
}

// Entry i: send box send[i] to rank peers[i] and receive recv[i] from it (the own rank copies directly).  Boxes are
// gathered into / scattered from compact staging with 3-D device copies; the transfers are one NCCL group.  This is
// the alltoallv of 2Decomp's transposes (deps/2Decomp_C/TransposeX2Y.cpp ... TransposeZ2Y.cpp) and, with one-row
// boxes, the y halo exchange.
int exchange_boxes(mifgpu_ctx *ctx, const std::vector<int> &peers, const std::vector<Box> &send, const std::vector<Box> &recv) {
  const int me = ctx->params.rank;
  std::vector<size_t> soff(peers.size()), roff(peers.size());
  size_t stotal = 0, rtotal = 0;
  for (size_t i = 0; i < peers.size(); i++) {
    soff[i] = stotal;
    roff[i] = rtotal;
    if (peers[i] == me) continue;
    stotal += send[i].count();
    rtotal += recv[i].count();
  }
  if (stotal > ctx->box_capacity || rtotal > ctx->box_capacity)
    return fail(MIFGPU_ERR_INVALID, "box exchange of %zu / %zu values exceeds the staging buffers (%zu)", stotal, rtotal, ctx->box_capacity);
  int rc;
  for (size_t i = 0; i < peers.size(); i++) {
    if (peers[i] == me) {
      if ((rc = copy_box(ctx, recv[i], send[i]))) return rc;
    } else if ((rc = copy_box(ctx, compact_box(ctx->box_send + soff[i], send[i]), send[i]))) {
      return rc;
    }
  }
  NCCL_TRY(g_nccl.GroupStart());
  for (size_t i = 0; i < peers.size(); i++) {
    if (peers[i] == me) continue;
    if (send[i].count()) NCCL_TRY(g_nccl.Send(ctx->box_send + soff[i], send[i].count(), kNcclReal, peers[i], ctx->comm, ctx->stream));
    if (recv[i].count()) NCCL_TRY(g_nccl.Recv(ctx->box_recv + roff[i], recv[i].count(), kNcclReal, peers[i], ctx->comm, ctx->stream));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  for (size_t i = 0; i < peers.size(); i++) {
    if (peers[i] == me) continue;
    if ((rc = copy_box(ctx, recv[i], compact_box(ctx->box_recv + roff[i], recv[i])))) return rc;
  }
  return MIFGPU_OK;
}

// 1-cell halo exchange in y of whole x-z sheets (src/StaggeredTensor.cpp:137-165): row 1 -> prev rank's last row, row
// sy-2 -> next rank's row 0.  Run BEFORE exchange_z, whose whole planes then carry the fresh y ghosts into the edge
// regions (SURVEY.md section 8a "edge ghosts": the result equals the single-rank one).
int exchange_y(mifgpu_ctx *ctx, mifgpu_tensor *const *tensors, int count, bool interior_only = false) {
  const Geom &g = ctx->g;
  if (g.prev_y == -1 && g.next_y == -1) return MIFGPU_OK;
  ProfScope prof(ctx, PROF_HALO);
  std::vector<int> peers;
  std::vector<Box> send, recv;
  for (int t = 0; t < count; t++) {
    const int s = tensors[t]->staggering;
    const int sy = g.sy[s];
    // interior_only: the part of a sheet the reference unpacks, 1 <= i <= sx - 2 and 1 <= k <= sz - 2
    // (src/StaggeredTensor.cpp:145-149,158-162)
    const int lo = interior_only ? 1 : 0, nx = g.sx[s] - 2 * lo, nz = g.sz[s] - 2 * lo;
    auto row = [&](int j) { return Box{tensors[t]->data, (size_t)g.PX, (size_t)g.PY, lo, j, lo, nx, 1, nz}; };
    if (g.prev_y != -1 && g.prev_y == g.next_y) {
      // periodic y on two y ranks: both neighbours are the same peer, and NCCL pairs the operations with one peer in issue
      // order (the reference tells them apart by tag): my row 1 is the peer's top ghost, my row sy - 2 its bottom ghost
      peers.push_back(g.prev_y);
      send.push_back(row(1));
      recv.push_back(row(sy - 1));
      peers.push_back(g.next_y);
      send.push_back(row(sy - 2));
      recv.push_back(row(0));
      continue;
    }
    if (g.prev_y != -1) {
      peers.push_back(g.prev_y);
      send.push_back(row(1));
      recv.push_back(row(0));
    }
    if (g.next_y != -1) {
      peers.push_back(g.next_y);
      send.push_back(row(sy - 2));
      recv.push_back(row(sy - 1));
    }
  }
  return exchange_boxes(ctx, peers, send, recv);
}

// The four 2Decomp transposes of the pencil decomposition as box exchanges.  which = 0: x pencil (owner region of
// `field`) -> y pencil; 1: y pencil -> z pencil; 2: z pencil -> y pencil; 3: y pencil -> x pencil.
int transpose_pencil(mifgpu_ctx *ctx, real *field, int which) {
  const Geom &g = ctx->g;
  ProfScope prof(ctx, PROF_TRANSPOSE);
  const int yr = ctx->y_rank, zr = ctx->z_rank, Py = ctx->Py, Pz = ctx->Pz;
  const int nxl = ctx->xs[yr + 1] - ctx->xs[yr], nyl = ctx->ys[yr + 1] - ctx->ys[yr], nzl = ctx->zs[zr + 1] - ctx->zs[zr];
  const int ny = ctx->ys[Py], nz = ctx->zs[Pz], nyz = ctx->yzs[zr + 1] - ctx->yzs[zr];
  const size_t pitch = (size_t)ctx->pen_pitch;
  std::vector<int> peers;
  std::vector<Box> a, b;  // a: boxes of the "earlier" layout, b: boxes of the "later" layout
  if (which == 0 || which == 3) {
    for (int r = 0; r < Py; r++) {
      peers.push_back(r * Pz + zr);
      // x pencil side: the x columns of y_rank r's pencils, my y rows, my z planes
      a.push_back(Box{field, (size_t)g.PX, (size_t)g.PY, g.own_lo[0] + ctx->xs[r], g.own_lo[1], g.own_lo[2],
                      ctx->xs[r + 1] - ctx->xs[r], nyl, nzl});
      // y pencil side: my x columns, the y rows of y_rank r, my z planes
      b.push_back(Box{ctx->ypen, pitch, (size_t)ny, 0, ctx->ys[r], 0, nxl, ctx->ys[r + 1] - ctx->ys[r], nzl});
    }
  } else {
    for (int r = 0; r < Pz; r++) {
      peers.push_back(yr * Pz + r);
      // y pencil side: my x columns, the y rows of z_rank r's z pencil, my z planes
      a.push_back(Box{ctx->ypen, pitch, (size_t)ny, 0, ctx->yzs[r], 0, nxl, ctx->yzs[r + 1] - ctx->yzs[r], nzl});
      // z pencil side: my x columns, my y rows, the z planes of z_rank r
      b.push_back(Box{ctx->zpen, pitch, (size_t)nyz, 0, 0, ctx->zs[r], nxl, nyz, ctx->zs[r + 1] - ctx->zs[r]});
    }
  }
  (void)nz;
  return (which == 0 || which == 1) ? exchange_boxes(ctx, peers, a, b) : exchange_boxes(ctx, peers, b, a);
}

int exchange_z(mifgpu_ctx *ctx, mifgpu_tensor *const *tensors, int count, cudaStream_t on = nullptr);

// Ghost refresh of `count` tensors in both split directions, y first (see exchange_y).
int exchange_halos(mifgpu_ctx *ctx, mifgpu_tensor *const *tensors, int count) {
  if (ctx->nranks == 1) return MIFGPU_OK;
  if (ctx->Py > 1 && ctx->reference_halos) {
    // The reference's own order and extents (src/StaggeredTensor.cpp:60-165): whole z planes are sent while their y
    // ghost rows still hold the previous exchange's values, and a received y sheet is unpacked without its i and k
    // borders -- so the y-ghost / z-ghost edges lag one exchange behind and a Py x Pz run differs from the one-rank run
    // (4e-8 in p after one step on 17^3, and the never-unpacked borders keep old values).  With this switch a Py x Pz run reproduces the reference's Py x Pz run instead.
    const int rc = exchange_z(ctx, tensors, count);
    if (rc) return rc;
    return exchange_y(ctx, tensors, count, true);
  }
  if (ctx->Py > 1) {
    const int rc = exchange_y(ctx, tensors, count);
    if (rc) return rc;
  }
  return exchange_z(ctx, tensors, count);
}

// Inside mifgpu_timestep the exchange that follows apply_bc (src/VelocityTensor.cpp:225-232) only has to deliver what
// is read before the next exchange of the same tensors (after the velocity correction, src/Timestep.cpp:68-75): the
// divergence reads w one plane above every owner point, i.e. w's ghost plane towards the next rank.  The correction
// touches owner planes only and its own exchange then refreshes every ghost plane of all three components, so the
// tensors end up exactly as with the reference's full exchange -- at one sixth of the traffic of this exchange.
int exchange_w_from_next(mifgpu_ctx *ctx, mifgpu_tensor *w, cudaStream_t on = nullptr) {
  if (ctx->nranks == 1) return MIFGPU_OK;
  const Geom &g = ctx->g;
  cudaStream_t stream = on ? on : ctx->stream;
  ProfScope prof(ctx, PROF_HALO, stream);
  const size_t plane = (size_t)g.plane;
  const int sz = g.sz[2];
  NCCL_TRY(g_nccl.GroupStart());
  if (g.prev_z != -1) NCCL_TRY(g_nccl.Send(w->data + plane, plane, kNcclReal, g.prev_z, ctx->comm, stream));
  if (g.next_z != -1) NCCL_TRY(g_nccl.Recv(w->data + plane * (sz - 1), plane, kNcclReal, g.next_z, ctx->comm, stream));
  NCCL_TRY(g_nccl.GroupEnd());
  return MIFGPU_OK;
}

// Start a z halo exchange behind everything queued on the compute stream so far and let the compute stream run on
// (overlap_halos; otherwise the exchange is simply queued on the compute stream).  halo_wait() orders the compute stream
// behind it.  w_top_only: the trimmed exchange that follows apply_bc inside the time step (exchange_w_from_next).
int halo_begin(mifgpu_ctx *ctx, mifgpu_tensor *const *tensors, int count, bool w_top_only) {
  if (ctx->nranks == 1) return MIFGPU_OK;
  if (!ctx->overlap_halos) return w_top_only ? exchange_w_from_next(ctx, tensors[0]) : exchange_z(ctx, tensors, count);
  CUDA_TRY(cudaEventRecord(ctx->ev_fork, ctx->stream));
  CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_fork, 0));
  const int rc = w_top_only ? exchange_w_from_next(ctx, tensors[0], ctx->comm_stream) : exchange_z(ctx, tensors, count, ctx->comm_stream);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->comm_stream));
  ctx->halo_pending = true;
  return MIFGPU_OK;
}
int halo_wait(mifgpu_ctx *ctx) {
  if (!ctx->halo_pending) return MIFGPU_OK;
  ctx->halo_pending = false;
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  return MIFGPU_OK;
}
// A stencil launch whose boundary planes read ghost planes that may still be in flight: interior planes [lo + inner_lo,
// hi - inner_hi) first, then the wait, then the boundary planes.  n = number of planes of the launcher's own range.
template <class Launch>
int launch_around_halo(mifgpu_ctx *ctx, int n, int inner_lo, int inner_hi, Launch launch) {
  if (!ctx->halo_pending || n - inner_lo - inner_hi <= 0) {
    const int rc = halo_wait(ctx);
    if (rc) return rc;
    launch(PlaneRange());
    return MIFGPU_OK;
  }
  PlaneRange interior, low, high;
  interior.first = inner_lo; interior.count = n - inner_lo - inner_hi;
  low.first = 0; low.count = inner_lo;
  high.first = n - inner_hi; high.count = inner_hi;
  launch(interior);
  const int rc = halo_wait(ctx);
  if (rc) return rc;
  if (inner_lo > 0) launch(low);
  if (inner_hi > 0) launch(high);
  return MIFGPU_OK;
}

int do_apply_bc(mifgpu_ctx *ctx, mifgpu_tensor *const vel[3], const mifgpu_bc *bc, double time, bool inside_timestep = false) {
  BcDev dev;
  std::memset(&dev, 0, sizeof(dev));
  dev.kind = bc->kind;
  dev.time = time;
  dev.Re = bc->Re;
  if (bc->kind == MIFGPU_BC_HOST_CALLBACK) {
    const int rc = fill_face_tables(ctx, bc, 0, time, time, dev);
    if (rc) return rc;
  } else if (bc->kind != MIFGPU_BC_TEST_CASE_1 && bc->kind != MIFGPU_BC_TEST_CASE_2 &&
             bc->kind != MIFGPU_BC_ETHIER_STEINMAN && bc->kind != MIFGPU_BC_VELOCITY_TEST) {
    return fail(MIFGPU_ERR_INVALID, "unknown boundary kind %d", bc->kind);
  }
  {
    ProfScope prof(ctx, PROF_BC);
    launch_apply_bc(ctx->stream, ctx->g, vec3(vel), dev, &ctx->launches);
  }
  static const bool full_exchange = getenv("MIFGPU_FULL_BC_EXCHANGE") != nullptr;  // A/B switch
  if (inside_timestep && !full_exchange && ctx->Py == 1) return halo_begin(ctx, &vel[2], 1, true);  // the divergence waits for it
  return exchange_halos(ctx, vel, 3);  // send_mpi_data / receive_mpi_data of the three components (src/VelocityTensor.cpp:225-232)
}

// solve_pressure_equation_homogeneous_periodic / _non_homogeneous_neumann (src/PressureEquation.cpp:266-286).
int do_solve(mifgpu_ctx *ctx, mifgpu_tensor *dp, mifgpu_tensor *const vel[3], double dt, const mifgpu_bc *nhn_bc,
             double t_new, double t_prev, bool leave_dp_exchange_pending = false) {
  {  // rhs = div(velocity) / dt (src/PressureEquation.cpp:59-61); only the top owner plane reads w's ghost plane
    ProfScope prof(ctx, PROF_DIVERGENCE);
    const int rc = launch_around_halo(ctx, ctx->g.own_hi[2] - ctx->g.own_lo[2], 0, 1, [&](PlaneRange planes) {
      launch_divergence(ctx->stream, ctx->g, cvec3(vel), 0.0, dt, dp->data, &ctx->launches, planes);
    });
    if (rc) return rc;
  }
  if (nhn_bc) {
    const Geom &g = ctx->g;
    if (g.periodic[0] || g.periodic[1] || g.periodic[2])
      return fail(MIFGPU_ERR_INVALID, "non-homogeneous Neumann data needs all directions non-periodic");
    BcDev dev;
    std::memset(&dev, 0, sizeof(dev));
    const int rc = fill_face_tables(ctx, nhn_bc, 1, t_new, t_prev, dev);
    if (rc) return rc;
    ProfScope prof(ctx, PROF_NHN);
    launch_nhn_rhs(ctx->stream, g, dp->data, dev, &ctx->launches);
  }
  // forward x, forward y, (forward z, eigenvalues, inverse z), inverse y, inverse x
  // (src/PressureEquation.cpp:79-101, 106-128, 133-195, 200-229, 234-263)
  auto sweep = [&](int dir, int mode, int category) {
    ProfScope prof(ctx, category);
    launch_poisson_sweep(ctx->stream, ctx->g, ctx->plan, dp->data, dir, mode, &ctx->launches);
  };
  if (ctx->Py > 1) {
    // Pencil decomposition: x sweeps on the sub-domain, y sweeps on the y pencil, the fused z sweep on the z pencil,
    // with 2Decomp's four transposes in between (src/PressureEquation.cpp:79-263).
    const int yr = ctx->y_rank, zr = ctx->z_rank;
    const int nxl = ctx->xs[yr + 1] - ctx->xs[yr], nzl = ctx->zs[zr + 1] - ctx->zs[zr];
    const int nyz = ctx->yzs[zr + 1] - ctx->yzs[zr];
    const int x0 = ctx->xs[yr], y0 = ctx->yzs[zr];
    int rc;
    sweep(0, 0, PROF_SWEEP_X_FWD);
    if ((rc = transpose_pencil(ctx, dp->data, 0))) return rc;
    {
      ProfScope prof(ctx, PROF_SWEEP_Y_FWD);
      launch_poisson_pencil(ctx->stream, ctx->plan, ctx->ypen, 1, 0, nxl, ctx->pen_pitch, nzl, x0, 0, false, &ctx->launches);
    }
    if ((rc = transpose_pencil(ctx, dp->data, 1))) return rc;
    {
      ProfScope prof(ctx, PROF_SWEEP_Z);
      launch_poisson_pencil(ctx->stream, ctx->plan, ctx->zpen, 2, 2, nxl, ctx->pen_pitch, nyz, x0, y0, x0 == 0 && y0 == 0,
                            &ctx->launches);
    }
    if ((rc = transpose_pencil(ctx, dp->data, 2))) return rc;
    {
      ProfScope prof(ctx, PROF_SWEEP_Y_INV);
      launch_poisson_pencil(ctx->stream, ctx->plan, ctx->ypen, 1, 1, nxl, ctx->pen_pitch, nzl, x0, 0, false, &ctx->launches);
    }
    if ((rc = transpose_pencil(ctx, dp->data, 3))) return rc;
    sweep(0, 1, PROF_SWEEP_X_INV);
    {
      ProfScope prof(ctx, PROF_PERIODIC);
      launch_periodic(ctx->stream, ctx->g, dp->data, 3, &ctx->launches);
    }
    mifgpu_tensor *one_pencil[1] = {dp};
    return exchange_halos(ctx, one_pencil, 1);
  }
  if (ctx->nranks > 1 && ctx->peer_mode && poisson_peer_capable(ctx->plan)) {
    // Transposes fused into the sweeps: the forward y sweep stores into the z pencils of the owning GPUs, the fused
    // z sweep stores into their slab staging buffers, the inverse y sweep reads its staging buffer (NVLink peer
    // stores; two stream-ordered rank barriers per solve).
    const PeerLayout peer = peer_layout(ctx);
    sweep(0, 0, PROF_SWEEP_X_FWD);
    int rc;
    {
      ProfScope prof(ctx, PROF_SWEEP_Y_FWD);
      launch_poisson_sweep_peer(ctx->stream, ctx->g, ctx->plan, dp->data, peer, 0, &ctx->launches);
      if ((rc = rank_barrier(ctx))) return rc;
    }
    {
      ProfScope prof(ctx, PROF_SWEEP_Z);
      launch_poisson_sweep_peer(ctx->stream, ctx->g, ctx->plan, dp->data, peer, 1, &ctx->launches);
      if ((rc = rank_barrier(ctx))) return rc;
    }
    {
      ProfScope prof(ctx, PROF_SWEEP_Y_INV);
      launch_poisson_sweep_peer(ctx->stream, ctx->g, ctx->plan, dp->data, peer, 2, &ctx->launches);
    }
    sweep(0, 1, PROF_SWEEP_X_INV);
    {
      ProfScope prof(ctx, PROF_PERIODIC);
      launch_periodic(ctx->stream, ctx->g, dp->data, 3, &ctx->launches);
    }
    mifgpu_tensor *one_peer[1] = {dp};
    const int hrc = halo_begin(ctx, one_peer, 1, false);
    return (hrc || leave_dp_exchange_pending) ? hrc : halo_wait(ctx);
  }
  sweep(0, 0, PROF_SWEEP_X_FWD);
  sweep(1, 0, PROF_SWEEP_Y_FWD);
  if (ctx->nranks == 1) {
    sweep(2, 2, PROF_SWEEP_Z);
  } else {
    int rc = transpose_slab(ctx, dp->data, true);
    if (rc) return rc;
    {
      ProfScope prof(ctx, PROF_SWEEP_Z);
      const int me = ctx->params.rank;
      launch_poisson_zpencil(ctx->stream, ctx->g, ctx->plan, ctx->zbuf, ctx->ylo[me + 1] - ctx->ylo[me], ctx->ylo[me],
                             ctx->ylo[me] == 0, &ctx->launches);
    }
    rc = transpose_slab(ctx, dp->data, false);
    if (rc) return rc;
  }
  sweep(1, 1, PROF_SWEEP_Y_INV);
  sweep(0, 1, PROF_SWEEP_X_INV);
  {
    ProfScope prof(ctx, PROF_PERIODIC);
    launch_periodic(ctx->stream, ctx->g, dp->data, 3, &ctx->launches);  // copy_to_staggered, src/PressureTensor.cpp:21-34
  }
  mifgpu_tensor *one[1] = {dp};
  const int hrc = halo_begin(ctx, one, 1, false);  // other.send_mpi_data(base_tag) / receive_mpi_data (src/PressureTensor.cpp:30-33)
  return (hrc || leave_dp_exchange_pending) ? hrc : halo_wait(ctx);
}

int check_launch(mifgpu_ctx *ctx) {
  (void)ctx;
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(MIFGPU_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(err));
  return MIFGPU_OK;
}

}  // namespace

extern "C" {

int mifgpu_abi_version(void) { return MIFGPU_ABI_VERSION; }

const char *mifgpu_last_error(void) { return g_last_error.c_str(); }
int mifgpu_real_bytes(void) { return (int)sizeof(real); }

static int create_context(const mifgpu_params *params, const void *unique_id, mifgpu_ctx **out);

int mifgpu_create(const mifgpu_params *params, mifgpu_ctx **out) {
  if (params && params->Py * params->Pz != 1)
    return fail(MIFGPU_ERR_INVALID, "Py*Pz = %d: use mifgpu_create_distributed with a communicator id", params->Py * params->Pz);
  return create_context(params, nullptr, out);
}

int mifgpu_comm_unique_id(void *unique_id) {
  if (!unique_id) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  static_assert(sizeof(ncclUniqueId) <= MIFGPU_UNIQUE_ID_BYTES, "unique id does not fit");
  if (!load_nccl()) return MIFGPU_ERR_COMM;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  std::memset(unique_id, 0, MIFGPU_UNIQUE_ID_BYTES);
  std::memcpy(unique_id, &id, sizeof(id));
  return MIFGPU_OK;
}

int mifgpu_create_distributed(const mifgpu_params *params, const void *unique_id, mifgpu_ctx **out) {
  if (!unique_id) return fail(MIFGPU_ERR_INVALID, "NULL communicator id");
  return create_context(params, unique_id, out);
}

static int create_context(const mifgpu_params *params, const void *unique_id, mifgpu_ctx **out) {
  if (!params || !out) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  *out = nullptr;
  Geom g;
  std::memset(&g, 0, sizeof(g));
  int n_points[3];
  int rc = build_geometry(*params, g, n_points);
  if (rc) return rc;
  if (params->Pz > 1 && (params->Nz_global - (params->periodic_bc[2] ? 1 : 0)) / params->Pz < 2)
    return fail(MIFGPU_ERR_INVALID, "fewer than 2 owner planes per rank");
  if (params->Py > 1 && (params->Ny_global / params->Py < 2 || (params->Nx_global - (params->periodic_bc[0] ? 1 : 0)) / params->Py < 1))
    return fail(MIFGPU_ERR_INVALID, "fewer than 2 owner rows (or no x column of the y pencil) per rank");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(MIFGPU_ERR_CUDA, "no CUDA device available (libmifgpu has no CPU fallback)");
  if (params->device < 0 || params->device >= count) return fail(MIFGPU_ERR_INVALID, "device %d out of range", params->device);
  CUDA_TRY(cudaSetDevice(params->device));
  mifgpu_ctx *ctx = new mifgpu_ctx();
  ctx->params = *params;
  ctx->g = g;
  for (int d = 0; d < 3; d++) ctx->n_points[d] = n_points[d];
  cudaError_t err = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (err != cudaSuccess) {
    delete ctx;
    return fail(MIFGPU_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(err));
  }
  const int per[3] = {g.periodic[0], g.periodic[1], g.periodic[2]};
  const double h[3] = {g.dx, g.dy, g.dz};
  const int n_global[3] = {(int)params->Nx_global, (int)params->Ny_global, (int)params->Nz_global};
  ctx->plan = poisson_plan_create(g, n_points, per, h, n_global);
  if (const int bad = poisson_plan_unsupported_direction(ctx->plan); bad >= 0) {
    mifgpu_destroy(ctx);
    return fail(MIFGPU_ERR_UNSUPPORTED, "%d transform points in direction %d: lines of more than 4097 points that are not 2^k + 1 "
                "do not fit the shared-memory transform of this build", n_points[bad], bad);
  }
  ctx->nranks = params->Py * params->Pz;
  ctx->Py = params->Py;
  ctx->Pz = params->Pz;
  ctx->y_rank = params->rank / params->Pz;
  ctx->z_rank = params->rank % params->Pz;
  ctx->reference_halos = getenv("MIFGPU_REFERENCE_HALOS") != nullptr;
  if (ctx->Py > 1) {
    // Pencil decomposition: block distributions of the transform points (src/Constants.cpp:78-79), NCCL communicator
    // over all Py * Pz ranks, pencil buffers and compact staging for the box exchanges.
    auto blocks = [](int n, int parts) {
      std::vector<int> first(parts + 1, 0);
      for (int r = 0; r < parts; r++) first[r + 1] = first[r] + n / parts + (r < n % parts ? 1 : 0);
      return first;
    };
    ctx->ys = blocks(n_points[1], ctx->Py);
    ctx->xs = blocks(n_points[0], ctx->Py);
    ctx->zs = blocks(n_points[2], ctx->Pz);
    ctx->yzs = blocks(n_points[1], ctx->Pz);
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    if (!load_nccl()) {
      mifgpu_destroy(ctx);
      return MIFGPU_ERR_COMM;
    }
    ncclResult_t nerr = g_nccl.CommInitRank(&ctx->comm, ctx->nranks, id, params->rank);
    if (nerr != ncclSuccess) {
      mifgpu_destroy(ctx);
      return fail(MIFGPU_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(nerr));
    }
    const int yr = ctx->y_rank, zr = ctx->z_rank;
    const size_t nxl = ctx->xs[yr + 1] - ctx->xs[yr], nzl = ctx->zs[zr + 1] - ctx->zs[zr], nyz = ctx->yzs[zr + 1] - ctx->yzs[zr];
    ctx->pen_pitch = (int)((nxl + 7) / 8 * 8);
    const size_t ypen = (size_t)ctx->pen_pitch * n_points[1] * nzl, zpen = (size_t)ctx->pen_pitch * nyz * n_points[2];
    // staging: the largest of one sub-domain, one pencil, and the y halo exchange of a velocity triple (three tensors, a
    // sheet towards each neighbour: 6 x-z sheets -- more than the whole sub-domain when a rank holds fewer than 6 rows)
    const size_t sub = (size_t)g.PX * g.PY * g.PZ;
    const size_t sheets = (size_t)6 * g.PX * g.PZ;
    const size_t stage = std::max(std::max(std::max(ypen, zpen), sub), sheets);
    ctx->box_capacity = stage;
    if (cudaMalloc(&ctx->ypen, ypen * sizeof(real)) != cudaSuccess || cudaMalloc(&ctx->zpen, zpen * sizeof(real)) != cudaSuccess ||
        cudaMalloc(&ctx->box_send, stage * sizeof(real)) != cudaSuccess || cudaMalloc(&ctx->box_recv, stage * sizeof(real)) != cudaSuccess) {
      mifgpu_destroy(ctx);
      return fail(MIFGPU_ERR_CUDA, "allocating the pencil buffers failed");
    }
    cudaMemset(ctx->ypen, 0, ypen * sizeof(real));
    cudaMemset(ctx->zpen, 0, zpen * sizeof(real));
  } else if (ctx->nranks > 1) {
    // MIF block distribution: the first (n mod P) ranks own one more point (src/Constants.cpp:78-79,
    // deps/2Decomp_C/C2Decomp.cpp:273-324).
    const int P = ctx->nranks;
    ctx->ylo.assign(P + 1, 0);
    ctx->zlo.assign(P + 1, 0);
    for (int r = 0; r < P; r++) {
      ctx->ylo[r + 1] = ctx->ylo[r] + n_points[1] / P + (r < n_points[1] % P ? 1 : 0);
      ctx->zlo[r + 1] = ctx->zlo[r] + n_points[2] / P + (r < n_points[2] % P ? 1 : 0);
    }
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    if (!load_nccl()) {
      mifgpu_destroy(ctx);
      return MIFGPU_ERR_COMM;
    }
    ncclResult_t nerr = g_nccl.CommInitRank(&ctx->comm, P, id, params->rank);
    if (nerr != ncclSuccess) {
      mifgpu_destroy(ctx);
      return fail(MIFGPU_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(nerr));
    }
    const int me = params->rank;
    const size_t slab = (size_t)(ctx->zlo[me + 1] - ctx->zlo[me]) * n_points[1] * g.PX;
    const size_t pencil = (size_t)n_points[2] * (ctx->ylo[me + 1] - ctx->ylo[me]) * g.PX;
    if (cudaMalloc(&ctx->xfer, slab * sizeof(real)) != cudaSuccess || cudaMalloc(&ctx->zbuf, pencil * sizeof(real)) != cudaSuccess ||
        cudaMalloc(&ctx->ylo_dev, (P + 1) * sizeof(int)) != cudaSuccess) {
      mifgpu_destroy(ctx);
      return fail(MIFGPU_ERR_CUDA, "allocating the transpose buffers failed");
    }
    cudaMemcpy(ctx->ylo_dev, ctx->ylo.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice);
    if (cudaMalloc(&ctx->barrier_word, 2 * sizeof(int)) != cudaSuccess || cudaMemset(ctx->barrier_word, 0, 2 * sizeof(int)) != cudaSuccess) {
      mifgpu_destroy(ctx);
      return fail(MIFGPU_ERR_CUDA, "allocating the barrier word failed");
    }
    if (P <= 8 && getenv("MIFGPU_NO_PEER") == nullptr) {
      const int prc = setup_peer_memory(ctx);
      if (prc) {
        mifgpu_destroy(ctx);
        return prc;
      }
    }
    if (getenv("MIFGPU_NO_HALO_OVERLAP") == nullptr) {  // A/B switch
      // highest priority: the exchange kernels get the first CTA slots the running stencil kernel frees
      int least = 0, greatest = 0;
      cudaDeviceGetStreamPriorityRange(&least, &greatest);
      if (cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, greatest) != cudaSuccess ||
          cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        mifgpu_destroy(ctx);
        return fail(MIFGPU_ERR_CUDA, "creating the communication stream failed");
      }
      ctx->overlap_halos = true;
    }
  }
  err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    mifgpu_destroy(ctx);
    return fail(MIFGPU_ERR_CUDA, "plan creation failed: %s", cudaGetErrorString(err));
  }
  *out = ctx;
  return MIFGPU_OK;
}

void mifgpu_destroy(mifgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->params.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int w = 0; w < 2; w++)
    for (int c = 0; c < 3; c++)
      for (int f = 0; f < 6; f++) {
        if (ctx->face_host[w][c][f]) cudaFreeHost(ctx->face_host[w][c][f]);
        if (ctx->face_dev[w][c][f]) cudaFree(ctx->face_dev[w][c][f]);
      }
  poisson_plan_destroy(ctx->plan);
  if (ctx->ypen) cudaFree(ctx->ypen);
  if (ctx->zpen) cudaFree(ctx->zpen);
  if (ctx->box_send) cudaFree(ctx->box_send);
  if (ctx->box_recv) cudaFree(ctx->box_recv);
  // Peer-mapped buffers: close this rank's mappings of the others, wait until every rank has done so, and only then
  // free the exported buffers (freeing memory a peer still has open through cudaIpcOpenMemHandle is undefined).
  for (int r = 0; r < 8; r++) {
    if (r == ctx->params.rank) continue;
    if (ctx->zbuf_peer[r]) cudaIpcCloseMemHandle(ctx->zbuf_peer[r]);
    if (ctx->xfer_peer[r]) cudaIpcCloseMemHandle(ctx->xfer_peer[r]);
  }
  if (ctx->peer_mode && ctx->comm && ctx->barrier_word) {
    g_nccl.AllReduce(ctx->barrier_word, ctx->barrier_word + 1, 1, ncclInt, ncclSum, ctx->comm, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  if (ctx->xfer) cudaFree(ctx->xfer);
  if (ctx->zbuf) cudaFree(ctx->zbuf);
  if (ctx->barrier_word) cudaFree(ctx->barrier_word);
  if (ctx->ylo_dev) cudaFree(ctx->ylo_dev);
  if (ctx->staging) cudaFree(ctx->staging);
  if (ctx->comm_stream) {
    cudaStreamSynchronize(ctx->comm_stream);
    cudaStreamDestroy(ctx->comm_stream);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  for (int d = 0; d < 4; d++) {
    if (ctx->copy_stream[d]) {
      cudaStreamSynchronize(ctx->copy_stream[d]);
      cudaStreamDestroy(ctx->copy_stream[d]);
    }
  }
  for (int d = 0; d < 2; d++)
    for (int b = 0; b < 2; b++) {
      if (ctx->copy_staging[d][b]) cudaFree(ctx->copy_staging[d][b]);
      if (ctx->ev_stage_filled[d][b]) cudaEventDestroy(ctx->ev_stage_filled[d][b]);
      if (ctx->ev_stage_free[d][b]) cudaEventDestroy(ctx->ev_stage_free[d][b]);
    }
  if (ctx->comm) g_nccl.CommDestroy(ctx->comm);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int mifgpu_tensor_extents(const mifgpu_ctx *ctx, int staggering, uint64_t extents[3]) {
  if (!ctx || !extents || staggering < 0 || staggering > 3) return fail(MIFGPU_ERR_INVALID, "bad argument");
  extents[0] = ctx->g.sx[staggering];
  extents[1] = ctx->g.sy[staggering];
  extents[2] = ctx->g.sz[staggering];
  return MIFGPU_OK;
}

int mifgpu_tensor_create(mifgpu_ctx *ctx, int staggering, mifgpu_tensor **out) {
  if (!ctx || !out || staggering < 0 || staggering > 3) return fail(MIFGPU_ERR_INVALID, "bad argument");
  *out = nullptr;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  real *data = nullptr;
  const size_t bytes = (size_t)ctx->g.volume * sizeof(real);
  CUDA_TRY(cudaMalloc(&data, bytes));
  cudaError_t err = cudaMemsetAsync(data, 0, bytes, ctx->stream);
  if (err != cudaSuccess) {
    cudaFree(data);
    return fail(MIFGPU_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(err));
  }
  mifgpu_tensor *t = new mifgpu_tensor();
  t->ctx = ctx;
  t->staggering = staggering;
  t->data = data;
  if (cudaEventCreateWithFlags(&t->ev_uploaded, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&t->ev_downloaded, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&t->ev_computed, cudaEventDisableTiming) != cudaSuccess) {
    mifgpu_tensor_destroy(t);
    return fail(MIFGPU_ERR_CUDA, "creating the tensor's events failed");
  }
  // the zero fill counts as the tensor's first "compute" use: an asynchronous transfer issued right after the creation
  // runs on another stream and must not overtake it
  if (cudaEventRecord(t->ev_computed, ctx->stream) != cudaSuccess) {
    mifgpu_tensor_destroy(t);
    return fail(MIFGPU_ERR_CUDA, "recording the tensor's creation failed");
  }
  *out = t;
  return MIFGPU_OK;
}

void mifgpu_tensor_destroy(mifgpu_tensor *t) {
  if (!t) return;
  cudaSetDevice(t->ctx->params.device);
  cudaStreamSynchronize(t->ctx->stream);
  for (cudaStream_t s : t->ctx->copy_stream)
    if (s) cudaStreamSynchronize(s);
  if (t->ev_uploaded) cudaEventDestroy(t->ev_uploaded);
  if (t->ev_downloaded) cudaEventDestroy(t->ev_downloaded);
  if (t->ev_computed) cudaEventDestroy(t->ev_computed);
  cudaFree(t->data);
  delete t;
}

// Host (compact reference extents) <-> device (padded pitches).  The PCIe transfer is one contiguous copy to or
// from a compact device staging buffer (row-pitched DMA over PCIe is several times slower); the re-pitching is a
// device-to-device 3-D copy.
static int copy_tensor(const mifgpu_tensor *t, real *host, bool to_device) {
  if (!t || !host) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  mifgpu_ctx *ctx = t->ctx;
  const Geom &g = ctx->g;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  const int s = t->staggering;
  const size_t bytes = (size_t)g.sx[s] * g.sy[s] * g.sz[s] * sizeof(real);
  if (ctx->staging_bytes < bytes) {
    if (ctx->staging) cudaFree(ctx->staging);
    ctx->staging = nullptr;
    ctx->staging_bytes = 0;
    CUDA_TRY(cudaMalloc(&ctx->staging, bytes));
    ctx->staging_bytes = bytes;
  }
  cudaMemcpy3DParms parms;
  std::memset(&parms, 0, sizeof(parms));
  const cudaPitchedPtr compact = make_cudaPitchedPtr(ctx->staging, (size_t)g.sx[s] * sizeof(real), g.sx[s], g.sy[s]);
  const cudaPitchedPtr padded = make_cudaPitchedPtr(t->data, (size_t)g.PX * sizeof(real), g.PX, g.PY);
  parms.srcPtr = to_device ? compact : padded;
  parms.dstPtr = to_device ? padded : compact;
  parms.extent = make_cudaExtent((size_t)g.sx[s] * sizeof(real), g.sy[s], g.sz[s]);
  parms.kind = cudaMemcpyDeviceToDevice;
  // asynchronous transfers of this tensor that are still in flight come first
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, t->ev_uploaded, 0));
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, t->ev_downloaded, 0));
  if (to_device) {
    CUDA_TRY(cudaMemcpyAsync(ctx->staging, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpy3DAsync(&parms, ctx->stream));
  } else {
    CUDA_TRY(cudaMemcpy3DAsync(&parms, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(host, ctx->staging, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_TRY(cudaEventRecord(t->ev_computed, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MIFGPU_OK;
}

// The same transfer on the copy stream of its direction, ordered by events only: after the last compute call that
// used the tensor and after its previous transfers; compute calls that use the tensor later wait for it (UseScope).
// Nothing blocks the host; `host` (pinned memory, or the copy degrades to a synchronous one) must stay untouched
// until mifgpu_synchronize.
static int copy_tensor_async(mifgpu_tensor *t, real *host, bool to_device) {
  if (!t || !host) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  mifgpu_ctx *ctx = t->ctx;
  const Geom &g = ctx->g;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  const int dir = to_device ? 0 : 1, s = t->staggering;
  for (int which : {dir, dir + 2})
    if (!ctx->copy_stream[which]) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream[which], cudaStreamNonBlocking));
  cudaStream_t link = ctx->copy_stream[dir], repack = ctx->copy_stream[dir + 2];
  const size_t bytes = (size_t)g.sx[s] * g.sy[s] * g.sz[s] * sizeof(real);
  if (ctx->copy_staging_bytes[dir] < bytes) {
    CUDA_TRY(cudaStreamSynchronize(link));
    CUDA_TRY(cudaStreamSynchronize(repack));
    // one size for all staggerings, so that the buffers are allocated once
    const size_t most = std::max(bytes, (size_t)(g.sx[0]) * (size_t)(g.sy[1]) * (size_t)(g.sz[2]) * sizeof(real));
    for (int b = 0; b < 2; b++) {
      if (ctx->copy_staging[dir][b]) cudaFree(ctx->copy_staging[dir][b]);
      ctx->copy_staging[dir][b] = nullptr;
      ctx->copy_staging_bytes[dir] = 0;
      CUDA_TRY(cudaMalloc(&ctx->copy_staging[dir][b], most));
      if (!ctx->ev_stage_filled[dir][b]) CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_stage_filled[dir][b], cudaEventDisableTiming));
      if (!ctx->ev_stage_free[dir][b]) CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_stage_free[dir][b], cudaEventDisableTiming));
    }
    ctx->copy_staging_bytes[dir] = most;
  }
  const int b = ctx->copy_next[dir];
  ctx->copy_next[dir] ^= 1;
  real *stage = ctx->copy_staging[dir][b];
  cudaMemcpy3DParms parms;
  std::memset(&parms, 0, sizeof(parms));
  const cudaPitchedPtr compact = make_cudaPitchedPtr(stage, (size_t)g.sx[s] * sizeof(real), g.sx[s], g.sy[s]);
  const cudaPitchedPtr padded = make_cudaPitchedPtr(t->data, (size_t)g.PX * sizeof(real), g.PX, g.PY);
  parms.srcPtr = to_device ? compact : padded;
  parms.dstPtr = to_device ? padded : compact;
  parms.extent = make_cudaExtent((size_t)g.sx[s] * sizeof(real), g.sy[s], g.sz[s]);
  parms.kind = cudaMemcpyDeviceToDevice;
  if (to_device) {
    // link: host -> stage b, once the re-pitching copy that last read stage b is done
    CUDA_TRY(cudaStreamWaitEvent(link, ctx->ev_stage_free[dir][b], 0));
    CUDA_TRY(cudaMemcpyAsync(stage, host, bytes, cudaMemcpyHostToDevice, link));
    CUDA_TRY(cudaEventRecord(ctx->ev_stage_filled[dir][b], link));
    // repack: stage b -> tensor, once the tensor is free (last compute call, last download of it)
    CUDA_TRY(cudaStreamWaitEvent(repack, ctx->ev_stage_filled[dir][b], 0));
    CUDA_TRY(cudaStreamWaitEvent(repack, t->ev_computed, 0));
    CUDA_TRY(cudaStreamWaitEvent(repack, t->ev_downloaded, 0));
    CUDA_TRY(cudaMemcpy3DAsync(&parms, repack));
    CUDA_TRY(cudaEventRecord(t->ev_uploaded, repack));
    CUDA_TRY(cudaEventRecord(ctx->ev_stage_free[dir][b], repack));
  } else {
    // repack: tensor -> stage b, once the tensor holds its values (last compute call, last upload) and the transfer
    // that last read stage b is done; the tensor may be overwritten as soon as this copy is done
    CUDA_TRY(cudaStreamWaitEvent(repack, t->ev_computed, 0));
    CUDA_TRY(cudaStreamWaitEvent(repack, t->ev_uploaded, 0));
    CUDA_TRY(cudaStreamWaitEvent(repack, ctx->ev_stage_free[dir][b], 0));
    CUDA_TRY(cudaMemcpy3DAsync(&parms, repack));
    CUDA_TRY(cudaEventRecord(ctx->ev_stage_filled[dir][b], repack));
    CUDA_TRY(cudaEventRecord(t->ev_downloaded, repack));
    // link: stage b -> host
    CUDA_TRY(cudaStreamWaitEvent(link, ctx->ev_stage_filled[dir][b], 0));
    CUDA_TRY(cudaMemcpyAsync(host, stage, bytes, cudaMemcpyDeviceToHost, link));
    CUDA_TRY(cudaEventRecord(ctx->ev_stage_free[dir][b], link));
  }
  return MIFGPU_OK;
}

int mifgpu_tensor_upload_async(mifgpu_tensor *t, const mifgpu_real *host) { return copy_tensor_async(t, const_cast<real *>(host), true); }
int mifgpu_tensor_download_async(mifgpu_tensor *t, mifgpu_real *host) { return copy_tensor_async(t, host, false); }

int mifgpu_tensor_upload(mifgpu_tensor *t, const mifgpu_real *host) { return copy_tensor(t, const_cast<real *>(host), true); }
int mifgpu_tensor_download(const mifgpu_tensor *t, mifgpu_real *host) { return copy_tensor(t, host, false); }

// Sub-box [lo, hi) of a tensor as one compact array: a device-to-device 3-D copy gathers the box from the padded
// tensor into the staging buffer (srcPos.x is in bytes, y / z in rows / slices), one contiguous copy crosses PCIe.
int mifgpu_tensor_download_box(const mifgpu_tensor *t, const int32_t lo[3], const int32_t hi[3], mifgpu_real *host) {
  if (!t || !lo || !hi || !host) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  mifgpu_ctx *ctx = t->ctx;
  const Geom &g = ctx->g;
  const int s = t->staggering;
  const int ext[3] = {g.sx[s], g.sy[s], g.sz[s]};
  for (int d = 0; d < 3; d++)
    if (lo[d] < 0 || hi[d] <= lo[d] || hi[d] > ext[d])
      return fail(MIFGPU_ERR_INVALID, "box [%d, %d) outside the tensor extent %d in direction %d", lo[d], hi[d], ext[d], d);
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  const size_t bx = (size_t)(hi[0] - lo[0]), by = (size_t)(hi[1] - lo[1]), bz = (size_t)(hi[2] - lo[2]);
  const size_t bytes = bx * by * bz * sizeof(real);
  if (ctx->staging_bytes < bytes) {
    if (ctx->staging) cudaFree(ctx->staging);
    ctx->staging = nullptr;
    ctx->staging_bytes = 0;
    CUDA_TRY(cudaMalloc(&ctx->staging, bytes));
    ctx->staging_bytes = bytes;
  }
  cudaMemcpy3DParms parms;
  std::memset(&parms, 0, sizeof(parms));
  parms.srcPtr = make_cudaPitchedPtr(t->data, (size_t)g.PX * sizeof(real), g.PX, g.PY);
  parms.srcPos = make_cudaPos((size_t)lo[0] * sizeof(real), (size_t)lo[1], (size_t)lo[2]);
  parms.dstPtr = make_cudaPitchedPtr(ctx->staging, bx * sizeof(real), bx, by);
  parms.dstPos = make_cudaPos(0, 0, 0);
  parms.extent = make_cudaExtent(bx * sizeof(real), by, bz);
  parms.kind = cudaMemcpyDeviceToDevice;
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, t->ev_uploaded, 0));
  CUDA_TRY(cudaMemcpy3DAsync(&parms, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(host, ctx->staging, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaEventRecord(t->ev_computed, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MIFGPU_OK;
}

int mifgpu_tensor_swap(mifgpu_tensor *a, mifgpu_tensor *b) {
  if (!a || !b || a->ctx != b->ctx || a->staggering != b->staggering) return fail(MIFGPU_ERR_INVALID, "tensors are not swappable");
  std::swap(a->data, b->data);
  std::swap(a->ev_uploaded, b->ev_uploaded);  // the events describe the device arrays, which just changed owners
  std::swap(a->ev_downloaded, b->ev_downloaded);
  std::swap(a->ev_computed, b->ev_computed);
  return MIFGPU_OK;
}

int mifgpu_apply_bc(mifgpu_ctx *ctx, mifgpu_tensor *const velocity[3], const mifgpu_bc *bc, double time) {
  if (!ctx || !bc) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  int rc = check_triple(ctx, velocity, "velocity");
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  UseScope use(ctx);
  use.add(velocity);
  rc = do_apply_bc(ctx, velocity, bc, time);
  if (rc) return rc;
  return check_launch(ctx);
}

int mifgpu_solve_pressure(mifgpu_ctx *ctx, mifgpu_tensor *pressure, mifgpu_tensor *const velocity[3], double dt,
                          const mifgpu_bc *nhn_bc, double nhn_time) {
  if (!ctx) return fail(MIFGPU_ERR_INVALID, "NULL context");
  int rc = check_triple(ctx, velocity, "velocity");
  if (rc) return rc;
  rc = check_scalar(ctx, pressure, "pressure");
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  UseScope use(ctx);
  use.add(velocity);
  use.add(pressure);
  rc = do_solve(ctx, pressure, velocity, dt, nhn_bc, nhn_time, nhn_time);
  if (rc) return rc;
  return check_launch(ctx);
}

int mifgpu_timestep(mifgpu_ctx *ctx, mifgpu_tensor *const velocity[3], mifgpu_tensor *const velocity_buffer[3],
                    mifgpu_tensor *const velocity_buffer_2[3], const mifgpu_bc *bc, double t_n, mifgpu_tensor *pressure,
                    mifgpu_tensor *pressure_buffer, int nhn) {
  if (!ctx || !bc) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  int rc;
  if ((rc = check_triple(ctx, velocity, "velocity"))) return rc;
  if ((rc = check_triple(ctx, velocity_buffer, "velocity_buffer"))) return rc;
  if ((rc = check_triple(ctx, velocity_buffer_2, "velocity_buffer_2"))) return rc;
  if ((rc = check_scalar(ctx, pressure, "pressure"))) return rc;
  if ((rc = check_scalar(ctx, pressure_buffer, "pressure_buffer"))) return rc;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  UseScope use(ctx);
  use.add(velocity);
  use.add(velocity_buffer);
  use.add(velocity_buffer_2);
  use.add(pressure);
  use.add(pressure_buffer);
  const Geom &g = ctx->g;
  cudaStream_t s = ctx->stream;
  // Stage times, src/Timestep.cpp:98-105.
  const double dt_1 = 64.0 / 120.0 * g.dt, time_1 = t_n + dt_1;
  const double dt_2 = 16.0 / 120.0 * g.dt, time_2 = time_1 + dt_2;
  const double dt_3 = 40.0 / 120.0 * g.dt, final_time = t_n + g.dt;
  const mifgpu_bc *nhn_bc = nhn ? bc : nullptr;

  // The stencil kernels of a stage read ghost planes only in their first / last planes, so each of them starts on its
  // interior planes while the halo exchange it depends on is still in flight (launch_around_halo); on one rank and with
  // pencils every launch covers all planes at once.
  const int nk_stage = std::max(g.sz[2], g.Nz) - 2;  // interior planes of the stage kernels: the lowest reads ghost plane 0,
                                                     // the highest the top ghost plane
  auto stage = [&](int category, int number, CVec3 in, Vec3 a, Vec3 b) {
    ProfScope prof(ctx, category);
    return launch_around_halo(ctx, nk_stage, 1, 1, [&](PlaneRange planes) {
      launch_stage(s, g, number, in, pressure->data, a, b, &ctx->launches, planes);
    });
  };
  // p += dp on all planes incl. ghosts and w -= dt grad_z(dp) at plane 1 read dp's ghost planes: planes 0, 1 and from Nz - 1
  // up wait for the exchange of dp; the velocity planes 1 and Nz - 2 that the next exchange sends are complete afterwards
  auto correct = [&](mifgpu_tensor *const vel[3], double dt_s) {
    ProfScope prof(ctx, PROF_CORRECT);
    return launch_around_halo(ctx, g.sz[2], 2, g.sz[2] - (g.Nz - 1), [&](PlaneRange planes) {
      launch_correct(s, g, vec3(vel), pressure->data, pressure_buffer->data, dt_s, &ctx->launches, planes);
    });
  };
  auto exchange_velocity = [&](mifgpu_tensor *const vel[3]) {  // src/Timestep.cpp:68-75
    if (ctx->Py > 1) return exchange_halos(ctx, vel, 3);
    return halo_begin(ctx, vel, 3, false);
  };

  // Stage 1 (src/Timestep.cpp:107-117).
  if ((rc = stage(PROF_STAGE1, 1, cvec3(velocity), vec3(velocity_buffer), vec3(velocity_buffer_2)))) return rc;
  if ((rc = do_apply_bc(ctx, velocity_buffer, bc, time_1, true))) return rc;
  if ((rc = do_solve(ctx, pressure_buffer, velocity_buffer, dt_1, nhn_bc, time_1, t_n, true))) return rc;
  if ((rc = correct(velocity_buffer, dt_1))) return rc;
  if ((rc = exchange_velocity(velocity_buffer))) return rc;

  // Stage 2 (src/Timestep.cpp:119-129).
  if ((rc = stage(PROF_STAGE2, 2, cvec3(velocity_buffer), vec3(velocity_buffer_2), vec3(velocity)))) return rc;
  if ((rc = do_apply_bc(ctx, velocity_buffer_2, bc, time_2, true))) return rc;
  if ((rc = do_solve(ctx, pressure_buffer, velocity_buffer_2, dt_2, nhn_bc, time_2, time_1, true))) return rc;
  if ((rc = correct(velocity_buffer_2, dt_2))) return rc;
  if ((rc = exchange_velocity(velocity_buffer_2))) return rc;

  // Stage 3 (src/Timestep.cpp:131-141).
  if ((rc = stage(PROF_STAGE3, 3, cvec3(velocity_buffer_2), vec3(velocity), vec3(velocity)))) return rc;
  if ((rc = do_apply_bc(ctx, velocity, bc, final_time, true))) return rc;
  if ((rc = do_solve(ctx, pressure_buffer, velocity, dt_3, nhn_bc, final_time, time_2, true))) return rc;
  if ((rc = correct(velocity, dt_3))) return rc;
  if ((rc = exchange_velocity(velocity))) return rc;
  if ((rc = halo_wait(ctx))) return rc;  // the step ends with complete ghost planes
  return check_launch(ctx);
}

int mifgpu_timestep_velocity(mifgpu_ctx *ctx, mifgpu_tensor *const velocity[3], mifgpu_tensor *const velocity_buffer[3],
                             mifgpu_tensor *const rhs_buffer[3], const mifgpu_bc *bc, double t_n) {
  if (!ctx || !bc) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  int rc;
  if ((rc = check_triple(ctx, velocity, "velocity"))) return rc;
  if ((rc = check_triple(ctx, velocity_buffer, "velocity_buffer"))) return rc;
  if ((rc = check_triple(ctx, rhs_buffer, "rhs_buffer"))) return rc;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  UseScope use(ctx);
  use.add(velocity);
  use.add(velocity_buffer);
  use.add(rhs_buffer);
  const Geom &g = ctx->g;
  cudaStream_t s = ctx->stream;
  // src/TimestepVelocity.cpp:11-13,60-63
  const double time_1 = t_n + (8.0 / 15.0) * g.dt, time_2 = t_n + (2.0 / 3.0) * g.dt, final_time = t_n + g.dt;
  {
    ProfScope prof(ctx, PROF_STAGE1);
    launch_velocity_stage(s, g, 1, cvec3(velocity), vec3(rhs_buffer), vec3(velocity_buffer), t_n, bc->Re, &ctx->launches);
  }
  if ((rc = do_apply_bc(ctx, velocity_buffer, bc, time_1))) return rc;
  {
    ProfScope prof(ctx, PROF_STAGE2);
    launch_velocity_stage(s, g, 2, cvec3(velocity_buffer), vec3(rhs_buffer), vec3(velocity), time_1, bc->Re, &ctx->launches);
  }
  if ((rc = do_apply_bc(ctx, velocity, bc, time_2))) return rc;
  {
    ProfScope prof(ctx, PROF_STAGE3);
    launch_velocity_stage(s, g, 3, cvec3(velocity), vec3(rhs_buffer), vec3(velocity_buffer), time_2, bc->Re, &ctx->launches);
  }
  if ((rc = do_apply_bc(ctx, velocity_buffer, bc, final_time))) return rc;
  for (int c = 0; c < 3; c++) std::swap(velocity[c]->data, velocity_buffer[c]->data);  // velocity.swap_data(velocity_buffer)
  return check_launch(ctx);
}

// ---- diagnostics on the device -----------------------------------------------------------------------------------
static int analytic_family(const mifgpu_bc *exact, double time, BcDev &dev) {
  if (!exact) return fail(MIFGPU_ERR_INVALID, "NULL analytic family");
  if (exact->kind != MIFGPU_BC_TEST_CASE_1 && exact->kind != MIFGPU_BC_TEST_CASE_2 && exact->kind != MIFGPU_BC_ETHIER_STEINMAN &&
      exact->kind != MIFGPU_BC_VELOCITY_TEST)
    return fail(MIFGPU_ERR_UNSUPPORTED, "device-side diagnostics need an analytic family evaluated on the device (kind %d given); "
                "download the tensors and use the host-side norms for arbitrary functions", exact->kind);
  std::memset(&dev, 0, sizeof(dev));
  dev.kind = exact->kind;
  dev.time = time;
  dev.Re = exact->Re;
  return MIFGPU_OK;
}

// Runs one of the error kernels and adds the per-CTA partial results up in CTA order: sums[0..3].
static int run_error_kernel(mifgpu_ctx *ctx, bool velocity, CVec3 vel, const real *p, const BcDev &dev, double sums[4]) {
  sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
  const int blocks = diag_blocks(ctx->g, velocity);
  if (blocks == 0) return MIFGPU_OK;
  real *partial = nullptr;
  CUDA_TRY(cudaMalloc(&partial, sizeof(real) * 4 * blocks));
  if (velocity) launch_velocity_error(ctx->stream, ctx->g, vel, dev, partial, &ctx->launches);
  else launch_pressure_error(ctx->stream, ctx->g, p, dev, partial, &ctx->launches);
  std::vector<real> host(4 * (size_t)blocks);
  cudaError_t err = cudaMemcpyAsync(host.data(), partial, sizeof(real) * host.size(), cudaMemcpyDeviceToHost, ctx->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
  cudaFree(partial);
  if (err != cudaSuccess) return fail(MIFGPU_ERR_CUDA, "error kernel failed: %s", cudaGetErrorString(err));
  for (int b = 0; b < blocks; b++) {
    sums[0] += host[4 * b];
    sums[1] += host[4 * b + 1];
    sums[2] = std::max(sums[2], (double)host[4 * b + 2]);
    sums[3] += host[4 * b + 3];
  }
  return MIFGPU_OK;
}

int mifgpu_velocity_error_norms(mifgpu_ctx *ctx, mifgpu_tensor *const velocity[3], const mifgpu_bc *exact, double time,
                                double norms[3]) {
  if (!ctx || !norms) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  int rc = check_triple(ctx, velocity, "velocity");
  if (rc) return rc;
  BcDev dev;
  if ((rc = analytic_family(exact, time, dev))) return rc;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  UseScope use(ctx);
  use.add(velocity);
  double sums[4];
  if ((rc = run_error_kernel(ctx, true, cvec3(velocity), nullptr, dev, sums))) return rc;
  const double cell = ctx->g.dx * ctx->g.dy * ctx->g.dz;
  norms[0] = sums[0] * cell;             // src/Norms.cpp:49-62
  norms[1] = std::sqrt(sums[1] * cell);  // src/Norms.cpp:64-76
  norms[2] = sums[2];                    // src/Norms.cpp:78-86
  return MIFGPU_OK;
}

int mifgpu_pressure_error_norms(mifgpu_ctx *ctx, const mifgpu_tensor *pressure, const mifgpu_bc *exact, double time,
                                double norms[3]) {
  if (!ctx || !norms) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  int rc = check_scalar(ctx, pressure, "pressure");
  if (rc) return rc;
  BcDev dev;
  if ((rc = analytic_family(exact, time, dev))) return rc;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  UseScope use(ctx);
  use.add(const_cast<mifgpu_tensor *>(pressure));
  double sums[4];
  if ((rc = run_error_kernel(ctx, false, CVec3{}, pressure->data, dev, sums))) return rc;
  const double cell = ctx->g.dx * ctx->g.dy * ctx->g.dz;
  norms[0] = sums[0] * cell;             // src/Norms.cpp:103-107
  norms[1] = std::sqrt(sums[1] * cell);  // src/Norms.cpp:109-113
  norms[2] = sums[2];                    // src/Norms.cpp:115-118
  return MIFGPU_OK;
}

int mifgpu_adjust_pressure(mifgpu_ctx *ctx, mifgpu_tensor *pressure, const mifgpu_bc *exact, double time) {
  if (!ctx) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  int rc = check_scalar(ctx, pressure, "pressure");
  if (rc) return rc;
  BcDev dev;
  if ((rc = analytic_family(exact, time, dev))) return rc;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  UseScope use(ctx);
  use.add(pressure);
  double sums[4];
  if ((rc = run_error_kernel(ctx, false, CVec3{}, pressure->data, dev, sums))) return rc;
  double difference = sums[3];
  if (ctx->nranks > 1) {
    // the rank-0 gather and broadcast of src/PressureEquation.cpp:298-333 as one all-reduce
    double *word = nullptr;
    CUDA_TRY(cudaMalloc(&word, sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(word, &difference, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(g_nccl.AllReduce(word, word, 1, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&difference, word, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaFree(word);
  }
  difference /= (double)ctx->n_points[0] * (double)ctx->n_points[1] * (double)ctx->n_points[2];  // :316-319
  launch_add_constant(ctx->stream, ctx->g, pressure->data, difference, &ctx->launches);
  return check_launch(ctx);
}

int mifgpu_rank_count(const mifgpu_ctx *ctx) { return ctx ? ctx->nranks : 0; }

int mifgpu_allreduce(mifgpu_ctx *ctx, double *values, int32_t count, int32_t op) {
  if (!ctx || !values || count < 0 || (op != 0 && op != 1)) return fail(MIFGPU_ERR_INVALID, "bad argument");
  if (ctx->nranks == 1 || count == 0) return MIFGPU_OK;
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  double *dev = nullptr;
  CUDA_TRY(cudaMalloc(&dev, sizeof(double) * count));
  CUDA_TRY(cudaMemcpyAsync(dev, values, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
  const ncclResult_t nerr = g_nccl.AllReduce(dev, dev, (size_t)count, ncclDouble, op == 0 ? ncclSum : ncclMax, ctx->comm, ctx->stream);
  cudaError_t err = cudaMemcpyAsync(values, dev, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
  cudaFree(dev);
  if (nerr != ncclSuccess) return fail(MIFGPU_ERR_COMM, "ncclAllReduce failed: %s", g_nccl.GetErrorString(nerr));
  if (err != cudaSuccess) return fail(MIFGPU_ERR_CUDA, "all-reduce copy failed: %s", cudaGetErrorString(err));
  return MIFGPU_OK;
}

int mifgpu_gather(mifgpu_ctx *ctx, const double *send, uint64_t count, double *recv, uint64_t *counts) {
  if (!ctx || !counts || (count > 0 && !send)) return fail(MIFGPU_ERR_INVALID, "bad argument");
  const int P = ctx->nranks, me = ctx->params.rank;
  std::vector<double> sizes(P, 0.0);
  sizes[me] = (double)count;
  int rc = mifgpu_allreduce(ctx, sizes.data(), P, 0);
  if (rc) return rc;
  uint64_t total = 0;
  for (int r = 0; r < P; r++) {
    counts[r] = (uint64_t)sizes[r];
    total += counts[r];
  }
  if (me == 0 && total > 0 && !recv) return fail(MIFGPU_ERR_INVALID, "rank 0 needs a receive buffer");
  if (P == 1) {
    if (count) std::memcpy(recv, send, sizeof(double) * count);
    return MIFGPU_OK;
  }
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  const uint64_t dev_count = (me == 0) ? total : count;
  if (dev_count == 0) return MIFGPU_OK;
  double *dev = nullptr;
  CUDA_TRY(cudaMalloc(&dev, sizeof(double) * dev_count));
  cudaError_t err = cudaSuccess;
  ncclResult_t nerr = ncclSuccess;
  if (count) err = cudaMemcpyAsync(dev, send, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream);  // rank 0: its own part comes first
  if (err == cudaSuccess) {
    nerr = g_nccl.GroupStart();
    if (me == 0) {
      uint64_t offset = counts[0];
      for (int r = 1; r < P && nerr == ncclSuccess; r++) {
        if (counts[r]) nerr = g_nccl.Recv(dev + offset, counts[r], ncclDouble, r, ctx->comm, ctx->stream);
        offset += counts[r];
      }
    } else if (count && nerr == ncclSuccess) {
      nerr = g_nccl.Send(dev, count, ncclDouble, 0, ctx->comm, ctx->stream);
    }
    const ncclResult_t end = g_nccl.GroupEnd();
    if (nerr == ncclSuccess) nerr = end;
  }
  if (err == cudaSuccess && nerr == ncclSuccess && me == 0)
    err = cudaMemcpyAsync(recv, dev, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
  cudaFree(dev);
  if (nerr != ncclSuccess) return fail(MIFGPU_ERR_COMM, "gather failed: %s", g_nccl.GetErrorString(nerr));
  if (err != cudaSuccess) return fail(MIFGPU_ERR_CUDA, "gather copy failed: %s", cudaGetErrorString(err));
  return MIFGPU_OK;
}

int mifgpu_slab_plan(uint64_t n_points, int32_t parts, int32_t *first) {
  if (!first || parts < 1) return fail(MIFGPU_ERR_INVALID, "bad argument");
  first[0] = 0;
  for (int r = 0; r < parts; r++) first[r + 1] = first[r] + (int32_t)(n_points / parts) + ((uint64_t)r < n_points % parts ? 1 : 0);
  return MIFGPU_OK;
}

int mifgpu_synchronize(mifgpu_ctx *ctx) {
  if (!ctx) return fail(MIFGPU_ERR_INVALID, "NULL context");
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ctx->comm_stream) CUDA_TRY(cudaStreamSynchronize(ctx->comm_stream));
  for (cudaStream_t s : ctx->copy_stream)
    if (s) CUDA_TRY(cudaStreamSynchronize(s));
  return MIFGPU_OK;
}

int mifgpu_transpose_path(const mifgpu_ctx *ctx) {
  if (!ctx) return -1;
  if (ctx->nranks == 1) return MIFGPU_TRANSPOSE_NONE;
  if (ctx->Py > 1) return MIFGPU_TRANSPOSE_PENCIL_BOXES;
  return (ctx->peer_mode && poisson_peer_capable(ctx->plan)) ? MIFGPU_TRANSPOSE_PEER_FUSED : MIFGPU_TRANSPOSE_NCCL_ALLTOALL;
}

int mifgpu_profile_enable(mifgpu_ctx *ctx, int enable) {
  if (!ctx) return fail(MIFGPU_ERR_INVALID, "NULL context");
  ctx->profiling = enable != 0;
  return MIFGPU_OK;
}

int mifgpu_profile_read(mifgpu_ctx *ctx, int capacity, const char **names, double *milliseconds, uint64_t *counts) {
  if (!ctx || !names || !milliseconds || !counts) return fail(MIFGPU_ERR_INVALID, "NULL argument");
  if (capacity < PROF_COUNT) return fail(MIFGPU_ERR_INVALID, "capacity must be >= %d", (int)PROF_COUNT);
  CUDA_TRY(cudaSetDevice(ctx->params.device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (int c = 0; c < PROF_COUNT; c++) {
    names[c] = kProfNames[c];
    milliseconds[c] = 0.0;
    counts[c] = 0;
  }
  for (ProfRecord &rec : ctx->prof_records) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, rec.start, rec.stop);
    milliseconds[rec.category] += ms;
    counts[rec.category] += 1;
    cudaEventDestroy(rec.start);
    cudaEventDestroy(rec.stop);
  }
  ctx->prof_records.clear();
  return PROF_COUNT;
}

void *mifgpu_stream(mifgpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

uint64_t mifgpu_launch_count(const mifgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
