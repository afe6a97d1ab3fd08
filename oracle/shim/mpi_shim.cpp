// oracle/shim/mpi_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Implementation of oracle/shim/mpi.h:
// thread-per-rank in-process MPI subset (see the header for the list of reference call sites covered).
#include "mpi.h"

#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

struct mifshim_comm {
  int ctx = 0;                 // context id, identical on every member (collective creation order)
  std::vector<int> members;    // world ranks, in communicator-rank order
  int ndims = 0;               // Cartesian topology (0 = none)
  int dims[3] = {1, 1, 1};
  int periods[3] = {0, 0, 0};
};

struct mifshim_file {
  int fd = -1;
};

namespace {

int g_size = 1;
thread_local int tl_rank = 0;
thread_local int tl_next_ctx = 2;  // 0 = world, 1 = self

mifshim_comm g_world;
thread_local mifshim_comm tl_self;

typedef std::tuple<int, int, int, int> Key;  // (ctx, src world rank, dst world rank, tag)
std::mutex g_mutex;
std::condition_variable g_cv;
std::map<Key, std::deque<std::vector<char>>> g_mailbox;

std::mutex g_type_mutex;
std::vector<size_t> g_type_sizes = {0, 1, 1, sizeof(int), sizeof(float), sizeof(double)};

size_t type_size(MPI_Datatype t) {
  std::lock_guard<std::mutex> lock(g_type_mutex);
  if (t <= 0 || static_cast<size_t>(t) >= g_type_sizes.size()) {
    std::fprintf(stderr, "mifshim: bad datatype handle %d\n", t);
    std::abort();
  }
  return g_type_sizes[t];
}

int comm_rank_of(const mifshim_comm *c) {
  for (size_t i = 0; i < c->members.size(); i++)
    if (c->members[i] == tl_rank) return static_cast<int>(i);
  std::fprintf(stderr, "mifshim: rank %d not in communicator\n", tl_rank);
  std::abort();
}

void post(const mifshim_comm *c, int dst_comm_rank, int tag, const void *buf, size_t bytes) {
  const int dst = c->members.at(dst_comm_rank);
  std::vector<char> msg(bytes);
  if (bytes) std::memcpy(msg.data(), buf, bytes);
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    g_mailbox[Key(c->ctx, tl_rank, dst, tag)].push_back(std::move(msg));
  }
  g_cv.notify_all();
}

void fetch(const mifshim_comm *c, int src_comm_rank, int tag, void *buf, size_t bytes) {
  const int src = c->members.at(src_comm_rank);
  const Key key(c->ctx, src, tl_rank, tag);
  std::unique_lock<std::mutex> lock(g_mutex);
  auto ready = [&] {
    auto it = g_mailbox.find(key);
    return it != g_mailbox.end() && !it->second.empty();
  };
  // A receive that stays unmatched for a long time is almost certainly a deadlock: say where.
  while (!g_cv.wait_for(lock, std::chrono::seconds(20), ready)) {
    std::fprintf(stderr, "mifshim: rank %d still waiting for (ctx %d, src %d, tag %d) after 20 s\n", tl_rank,
                 c->ctx, src, tag);
    if (std::getenv("MIF_SHIM_ABORT_ON_STALL")) std::abort();
  }
  auto &queue = g_mailbox[key];
  std::vector<char> msg = std::move(queue.front());
  queue.pop_front();
  lock.unlock();
  if (msg.size() > bytes) {
    std::fprintf(stderr, "mifshim: message truncated (%zu > %zu bytes, tag %d)\n", msg.size(), bytes, tag);
    std::abort();
  }
  if (!msg.empty()) std::memcpy(buf, msg.data(), msg.size());
}

// Internal tags for collectives (user tags are >= 0).
enum { TAG_BARRIER = -10, TAG_BCAST = -11, TAG_GATHER = -12, TAG_ALLGATHER = -13, TAG_ALLREDUCE = -14,
       TAG_ALLTOALL = -15 };

}  // namespace

MPI_Comm mifshim_comm_world() { return &g_world; }

MPI_Comm mifshim_comm_self() {
  tl_self.ctx = 1;
  tl_self.members.assign(1, tl_rank);
  return &tl_self;
}

int mifshim_run(int np, int (*fn)(int, char **), int argc, char **argv) {
  g_size = np;
  g_world.ctx = 0;
  g_world.members.resize(np);
  for (int r = 0; r < np; r++) g_world.members[r] = r;
  std::vector<int> rc(np, 0);
  if (np == 1) {
    tl_rank = 0;
    return fn(argc, argv);
  }
  std::vector<std::thread> threads;
  for (int r = 0; r < np; r++) {
    threads.emplace_back([&, r] {
      tl_rank = r;
      rc[r] = fn(argc, argv);
    });
  }
  int worst = 0;
  for (int r = 0; r < np; r++) {
    threads[r].join();
    if (rc[r] > worst) worst = rc[r];
  }
  return worst;
}

int MPI_Init(int *, char ***) {
  if (g_world.members.empty()) {  // not launched through mifshim_run: single rank
    g_size = 1;
    g_world.members.assign(1, 0);
  }
  return MPI_SUCCESS;
}
int MPI_Finalize() { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int errorcode) {
  std::fprintf(stderr, "mifshim: MPI_Abort(%d)\n", errorcode);
  std::_Exit(errorcode ? errorcode : 1);
}
double MPI_Wtime() {
  using clock = std::chrono::steady_clock;
  return std::chrono::duration<double>(clock::now().time_since_epoch()).count();
}
int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  if (comm->members.empty() && comm == &g_world) MPI_Init(nullptr, nullptr);
  *rank = comm_rank_of(comm);
  return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm comm, int *size) {
  if (comm->members.empty() && comm == &g_world) MPI_Init(nullptr, nullptr);
  *size = static_cast<int>(comm->members.size());
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm) {
  const int n = static_cast<int>(comm->members.size());
  const int me = comm_rank_of(comm);
  char token = 0;
  if (me == 0) {
    for (int r = 1; r < n; r++) fetch(comm, r, TAG_BARRIER, &token, 1);
    for (int r = 1; r < n; r++) post(comm, r, TAG_BARRIER, &token, 1);
  } else {
    post(comm, 0, TAG_BARRIER, &token, 1);
    fetch(comm, 0, TAG_BARRIER, &token, 1);
  }
  return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype) {
  const size_t bytes = type_size(oldtype) * static_cast<size_t>(count);
  std::lock_guard<std::mutex> lock(g_type_mutex);
  g_type_sizes.push_back(bytes);
  *newtype = static_cast<int>(g_type_sizes.size()) - 1;
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *) { return MPI_SUCCESS; }
int MPI_Type_size(MPI_Datatype type, int *size) {
  *size = static_cast<int>(type_size(type));
  return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
  if (dest == MPI_PROC_NULL) return MPI_SUCCESS;
  post(comm, dest, tag, buf, type_size(type) * static_cast<size_t>(count));
  return MPI_SUCCESS;
}
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm,
              MPI_Request *req) {
  MPI_Send(buf, count, type, dest, tag, comm);  // eager: the payload is copied, the request is complete
  *req = 1;
  return MPI_SUCCESS;
}
int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status *status) {
  if (source == MPI_PROC_NULL) return MPI_SUCCESS;
  fetch(comm, source, tag, buf, type_size(type) * static_cast<size_t>(count));
  if (status) {
    status->MPI_SOURCE = source;
    status->MPI_TAG = tag;
    status->MPI_ERROR = MPI_SUCCESS;
  }
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *req, MPI_Status *) {
  *req = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request *reqs, MPI_Status *) {
  for (int i = 0; i < count; i++) reqs[i] = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
  const int n = static_cast<int>(comm->members.size());
  const int me = comm_rank_of(comm);
  const size_t bytes = type_size(type) * static_cast<size_t>(count);
  if (me == root) {
    for (int r = 0; r < n; r++)
      if (r != root) post(comm, r, TAG_BCAST, buf, bytes);
  } else {
    fetch(comm, root, TAG_BCAST, buf, bytes);
  }
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf, const int *rcounts,
                const int *displs, MPI_Datatype rtype, int root, MPI_Comm comm) {
  const int n = static_cast<int>(comm->members.size());
  const int me = comm_rank_of(comm);
  const size_t sbytes = type_size(stype) * static_cast<size_t>(scount);
  // Zero-sized contributions are neither sent nor awaited, like the linear gatherv of MPICH/Open MPI:
  // the reference's writeDat (src/VTKDatExport.cpp:540) lets ranks without points skip the call.
  if (me != root) {
    if (sbytes) post(comm, root, TAG_GATHER, sbuf, sbytes);
    return MPI_SUCCESS;
  }
  const size_t rsz = type_size(rtype);
  for (int r = 0; r < n; r++) {
    char *dst = static_cast<char *>(rbuf) + rsz * static_cast<size_t>(displs[r]);
    const size_t bytes = rsz * static_cast<size_t>(rcounts[r]);
    if (r == me) {
      if (sbytes) std::memcpy(dst, sbuf, sbytes);
    } else if (bytes) {
      fetch(comm, r, TAG_GATHER, dst, bytes);
    }
  }
  return MPI_SUCCESS;
}

int MPI_Gather(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf, int rcount, MPI_Datatype rtype,
               int root, MPI_Comm comm) {
  const int n = static_cast<int>(comm->members.size());
  std::vector<int> counts(n, rcount), displs(n);
  for (int r = 0; r < n; r++) displs[r] = r * rcount;
  return MPI_Gatherv(sbuf, scount, stype, rbuf, counts.data(), displs.data(), rtype, root, comm);
}

int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf, int rcount,
                  MPI_Datatype rtype, MPI_Comm comm) {
  const int n = static_cast<int>(comm->members.size());
  const int me = comm_rank_of(comm);
  const size_t sbytes = type_size(stype) * static_cast<size_t>(scount);
  const size_t rbytes = type_size(rtype) * static_cast<size_t>(rcount);
  for (int r = 0; r < n; r++)
    if (r != me) post(comm, r, TAG_ALLGATHER, sbuf, sbytes);
  for (int r = 0; r < n; r++) {
    char *dst = static_cast<char *>(rbuf) + rbytes * static_cast<size_t>(r);
    if (r == me) std::memcpy(dst, sbuf, sbytes);
    else fetch(comm, r, TAG_ALLGATHER, dst, rbytes);
  }
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
  const int n = static_cast<int>(comm->members.size());
  const int me = comm_rank_of(comm);
  const size_t bytes = type_size(type) * static_cast<size_t>(count);
  std::vector<char> all(bytes * n);
  for (int r = 0; r < n; r++)
    if (r != me) post(comm, r, TAG_ALLREDUCE, sbuf, bytes);
  for (int r = 0; r < n; r++) {
    if (r == me) std::memcpy(all.data() + bytes * r, sbuf, bytes);
    else fetch(comm, r, TAG_ALLREDUCE, all.data() + bytes * r, bytes);
  }
  auto reduce = [&](auto *out) {
    typedef typename std::remove_reference<decltype(*out)>::type T;
    for (int i = 0; i < count; i++) {
      T acc = reinterpret_cast<const T *>(all.data())[i];
      for (int r = 1; r < n; r++) {
        const T v = reinterpret_cast<const T *>(all.data() + bytes * r)[i];
        if (op == MPI_SUM) acc += v;
        else if (op == MPI_MAX) acc = v > acc ? v : acc;
        else acc = v < acc ? v : acc;
      }
      out[i] = acc;
    }
  };
  if (type == MPI_DOUBLE) reduce(static_cast<double *>(rbuf));
  else if (type == MPI_FLOAT) reduce(static_cast<float *>(rbuf));
  else if (type == MPI_INT) reduce(static_cast<int *>(rbuf));
  else {
    std::fprintf(stderr, "mifshim: MPI_Allreduce on unsupported datatype %d\n", type);
    std::abort();
  }
  return MPI_SUCCESS;
}

int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype stype, void *rbuf,
                  const int *rcounts, const int *rdispls, MPI_Datatype rtype, MPI_Comm comm) {
  const int n = static_cast<int>(comm->members.size());
  const int me = comm_rank_of(comm);
  const size_t ssz = type_size(stype), rsz = type_size(rtype);
  // The self block is staged through a copy because 2Decomp passes overlapping work buffers only
  // between distinct arrays, but a memmove keeps this safe in any case.
  for (int r = 0; r < n; r++) {
    const char *src = static_cast<const char *>(sbuf) + ssz * static_cast<size_t>(sdispls[r]);
    if (r != me) post(comm, r, TAG_ALLTOALL, src, ssz * static_cast<size_t>(scounts[r]));
  }
  {
    const char *src = static_cast<const char *>(sbuf) + ssz * static_cast<size_t>(sdispls[me]);
    char *dst = static_cast<char *>(rbuf) + rsz * static_cast<size_t>(rdispls[me]);
    std::memmove(dst, src, ssz * static_cast<size_t>(scounts[me]));
  }
  for (int r = 0; r < n; r++) {
    char *dst = static_cast<char *>(rbuf) + rsz * static_cast<size_t>(rdispls[r]);
    if (r != me) fetch(comm, r, TAG_ALLTOALL, dst, rsz * static_cast<size_t>(rcounts[r]));
  }
  return MPI_SUCCESS;
}

int MPI_Ialltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype stype, void *rbuf,
                   const int *rcounts, const int *rdispls, MPI_Datatype rtype, MPI_Comm comm,
                   MPI_Request *req) {
  *req = 1;
  return MPI_Alltoallv(sbuf, scounts, sdispls, stype, rbuf, rcounts, rdispls, rtype, comm);
}

int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int, MPI_Comm *newcomm) {
  mifshim_comm *c = new mifshim_comm(*comm);  // same members, row-major rank order (MPI standard)
  c->ctx = tl_next_ctx++;
  c->ndims = ndims;
  size_t total = 1;
  for (int d = 0; d < ndims; d++) {
    c->dims[d] = dims[d];
    c->periods[d] = periods[d];
    total *= static_cast<size_t>(dims[d]);
  }
  if (total != c->members.size()) {
    std::fprintf(stderr, "mifshim: MPI_Cart_create dims do not match communicator size\n");
    std::abort();
  }
  *newcomm = c;
  return MPI_SUCCESS;
}

int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords) {
  int rem = rank;
  for (int d = comm->ndims - 1; d >= 0; d--) {
    if (d < maxdims) coords[d] = rem % comm->dims[d];
    rem /= comm->dims[d];
  }
  return MPI_SUCCESS;
}

int MPI_Cart_sub(MPI_Comm comm, const int *remain_dims, MPI_Comm *newcomm) {
  const int me = comm_rank_of(comm);
  int my_coords[3] = {0, 0, 0};
  MPI_Cart_coords(comm, me, comm->ndims, my_coords);
  mifshim_comm *c = new mifshim_comm();
  c->ctx = tl_next_ctx++;
  c->ndims = 0;
  for (int d = 0; d < comm->ndims; d++) {
    if (remain_dims[d]) {
      c->dims[c->ndims] = comm->dims[d];
      c->periods[c->ndims] = comm->periods[d];
      c->ndims++;
    }
  }
  // Members: all ranks of `comm` sharing my coordinates in the dropped dimensions, in rank order.
  // Sub-communicators of different groups get the same ctx, which is harmless: the mailbox key also
  // contains the (disjoint) world ranks.
  for (int r = 0; r < static_cast<int>(comm->members.size()); r++) {
    int coords[3] = {0, 0, 0};
    MPI_Cart_coords(comm, r, comm->ndims, coords);
    bool same = true;
    for (int d = 0; d < comm->ndims; d++)
      if (!remain_dims[d] && coords[d] != my_coords[d]) same = false;
    if (same) c->members.push_back(comm->members[r]);
  }
  *newcomm = c;
  return MPI_SUCCESS;
}

int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest) {
  const int me = comm_rank_of(comm);
  int coords[3] = {0, 0, 0};
  MPI_Cart_coords(comm, me, comm->ndims, coords);
  auto neighbour = [&](int delta) -> int {
    int c[3] = {coords[0], coords[1], coords[2]};
    c[direction] += delta;
    const int n = comm->dims[direction];
    if (c[direction] < 0 || c[direction] >= n) {
      if (!comm->periods[direction]) return MPI_PROC_NULL;
      c[direction] = ((c[direction] % n) + n) % n;
    }
    int rank = 0;
    for (int d = 0; d < comm->ndims; d++) rank = rank * comm->dims[d] + c[d];
    return rank;
  };
  *rank_source = neighbour(-disp);
  *rank_dest = neighbour(disp);
  return MPI_SUCCESS;
}

int MPI_File_open(MPI_Comm, const char *filename, int amode, MPI_Info, MPI_File *fh) {
  int flags = 0;
  if (amode & MPI_MODE_RDWR) flags |= O_RDWR;
  else if (amode & MPI_MODE_WRONLY) flags |= O_WRONLY;
  else flags |= O_RDONLY;
  if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
  const int fd = ::open(filename, flags, 0644);
  if (fd < 0) return MPI_ERR_OTHER;
  *fh = new mifshim_file();
  (*fh)->fd = fd;
  return MPI_SUCCESS;
}
int MPI_File_close(MPI_File *fh) {
  if (fh && *fh) {
    ::close((*fh)->fd);
    delete *fh;
    *fh = nullptr;
  }
  return MPI_SUCCESS;
}
int MPI_File_delete(const char *filename, MPI_Info) {
  return ::unlink(filename) == 0 ? MPI_SUCCESS : MPI_ERR_OTHER;
}
int MPI_File_write_at(MPI_File fh, MPI_Offset offset, const void *buf, int count, MPI_Datatype type,
                      MPI_Status *) {
  const size_t bytes = type_size(type) * static_cast<size_t>(count);
  size_t done = 0;
  while (done < bytes) {
    const ssize_t w = ::pwrite(fh->fd, static_cast<const char *>(buf) + done, bytes - done,
                               static_cast<off_t>(offset) + static_cast<off_t>(done));
    if (w <= 0) return MPI_ERR_OTHER;
    done += static_cast<size_t>(w);
  }
  return MPI_SUCCESS;
}
