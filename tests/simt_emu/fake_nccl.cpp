// tests/simt_emu/fake_nccl.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The nine NCCL entry points libmifgpu binds with dlopen (csrc/mif_api.cu, load_nccl), implemented between PROCESSES
// of one machine through mailbox files in /dev/shm, so that the multi-rank paths of the library (plane halos, the
// grouped send/recv all-to-all, the rank barriers of the peer-memory transposes) can run under the SIMT interpreter
// on a machine without GPUs.  Everything is synchronous: a send writes its payload to a file named after
// (communicator, source, destination, sequence number) and renames it into place; a receive polls for the file of
// the next sequence number of that pair.  Sends never block; receives issued between ncclGroupStart and ncclGroupEnd are
// queued and carried out at ncclGroupEnd, after all sends of the group -- like NCCL, where nothing of a group runs before
// it is closed -- so rings of three or more ranks (periodic z) cannot deadlock on a receive that precedes a send.
//
// MIF_FAKE_NCCL_LATE=1 models the other extreme of an exchange that runs on a second stream next to the compute stream
// (the overlapped halo exchanges of mif_api.cu): every send / receive issued on a stream that was created with a priority
// is carried out as LATE as stream order allows -- when the interpreter's runtime reports that another
// stream waits for an event recorded behind it, or that the stream is synchronised (fake_nccl_flush, called from
// emu_runtime.cpp).  A kernel that reads a ghost plane, or overwrites a plane still to be sent, before that wait then
// fails the parity tests; the default mode (everything at once) covers the earliest possible completion.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <utility>
#include <vector>

extern "C" {

typedef struct FakeComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;  // ncclChar = 0, ncclInt = 2, ncclDouble = 8 (nccl.h)
typedef int ncclRedOp_t;
typedef void *cudaStream_t;

struct FakeComm {
  std::string tag;
  int rank, nranks;
  std::vector<uint64_t> sent, received;  // per peer sequence numbers
};

static size_t type_size(ncclDataType_t t) {
  switch (t) {
    case 0: case 1: return 1;
    case 2: case 3: return 4;
    case 4: case 5: case 8: return 8;
    case 7: return 4;
    default: return 1;
  }
}

static std::string mailbox(const FakeComm *c, int src, int dst, uint64_t seq) {
  char name[256];
  std::snprintf(name, sizeof(name), "/dev/shm/mif_fake_nccl_%s_%d_%d_%llu", c->tag.c_str(), src, dst, (unsigned long long)seq);
  return name;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t c,
                           cudaStream_t s);

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  std::memset(id, 0, sizeof(*id));
  std::snprintf(id->internal, sizeof(id->internal), "%d_%ld", (int)getpid(), (long)random());
  return 0;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  FakeComm *c = new FakeComm();
  id.internal[127] = 0;
  c->tag = id.internal;
  c->rank = rank;
  c->nranks = nranks;
  c->sent.assign(nranks, 0);
  c->received.assign(nranks, 0);
  *comm = c;
  // like the real call, return only after every rank has joined
  int one = 1, sum = 0;
  return ncclAllReduce(&one, &sum, 1, 2 /* ncclInt */, 0 /* ncclSum */, c, nullptr);
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  delete comm;
  return 0;
}

struct QueuedRecv { void *buf; size_t count; ncclDataType_t type; int peer; ncclComm_t comm; };
static int g_group_depth = 0;
static std::vector<QueuedRecv> g_queued;
static ncclResult_t recv_now(void *buf, size_t count, ncclDataType_t type, int peer, ncclComm_t c);

static ncclResult_t send_now(const void *buf, size_t count, ncclDataType_t type, int peer, ncclComm_t c);

// The interpreter's runtime registers its "carry out what is queued on this stream" here (MIF_EMU_LAZY_COPIES): an
// operation issued on a stream runs behind the asynchronous copies queued on that stream before it.
static void (*g_stream_flush)(void *) = nullptr;
void fake_nccl_set_stream_flush(void (*flush)(void *)) { g_stream_flush = flush; }
static void behind_stream(cudaStream_t s) {
  if (g_stream_flush) g_stream_flush(s);
}

// late mode: operations of side streams wait here, in issue order (sends before the receives of their group)
struct LateOp { cudaStream_t stream; bool is_send; std::function<ncclResult_t()> run; };
static std::vector<LateOp> g_late;
static bool is_late(cudaStream_t s) {
  static const bool enabled = getenv("MIF_FAKE_NCCL_LATE") != nullptr;
  // emu_runtime.cpp's streams start with an int that is 1 for streams created with a priority (the side streams)
  return enabled && s != nullptr && *static_cast<const int *>(s) == 1;
}
// carry out what was deferred on `stream` (nullptr: on every stream)
int fake_nccl_flush(cudaStream_t stream) {
  int rc = 0;
  std::vector<LateOp> keep, run;
  for (auto &op : g_late) (stream == nullptr || op.stream == stream ? run : keep).push_back(std::move(op));
  g_late.swap(keep);
  // all sends first, then the receives (each kind in issue order): ranks flushing at the same point cannot block
  for (auto &op : run) if (op.is_send) rc |= op.run();
  for (auto &op : run) if (!op.is_send) rc |= op.run();
  return rc;
}

ncclResult_t ncclGroupStart() {
  g_group_depth++;
  return 0;
}
ncclResult_t ncclGroupEnd() {
  if (g_group_depth > 0) g_group_depth--;
  if (g_group_depth > 0) return 0;
  ncclResult_t rc = 0;
  for (const QueuedRecv &q : g_queued)
    if (recv_now(q.buf, q.count, q.type, q.peer, q.comm)) rc = 1;
  g_queued.clear();
  return rc;
}

ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t type, int peer, ncclComm_t c, cudaStream_t s) {
  behind_stream(s);
  if (is_late(s)) {
    g_late.push_back(LateOp{s, true, [=] { return send_now(buf, count, type, peer, c); }});
    return 0;
  }
  return send_now(buf, count, type, peer, c);
}

static ncclResult_t send_now(const void *buf, size_t count, ncclDataType_t type, int peer, ncclComm_t c) {
  const size_t bytes = count * type_size(type);
  const std::string final_name = mailbox(c, c->rank, peer, c->sent[peer]++);
  const std::string tmp = final_name + ".tmp";
  FILE *f = std::fopen(tmp.c_str(), "wb");
  if (!f) return 1;
  const size_t written = bytes ? std::fwrite(buf, 1, bytes, f) : 0;
  std::fclose(f);
  if (written != bytes || std::rename(tmp.c_str(), final_name.c_str()) != 0) return 1;
  return 0;
}

ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t type, int peer, ncclComm_t c, cudaStream_t s) {
  behind_stream(s);
  if (is_late(s)) {
    g_late.push_back(LateOp{s, false, [=] { return recv_now(buf, count, type, peer, c); }});
    return 0;
  }
  if (g_group_depth > 0) {
    g_queued.push_back(QueuedRecv{buf, count, type, peer, c});
    return 0;
  }
  return recv_now(buf, count, type, peer, c);
}

static ncclResult_t recv_now(void *buf, size_t count, ncclDataType_t type, int peer, ncclComm_t c) {
  const size_t bytes = count * type_size(type);
  const std::string name = mailbox(c, peer, c->rank, c->received[peer]++);
  for (long spins = 0;; spins++) {
    struct stat st;
    if (stat(name.c_str(), &st) == 0) break;
    if (spins > 600000) {  // 10 minutes
      std::fprintf(stderr, "fake_nccl: rank %d timed out waiting for %s\n", c->rank, name.c_str());
      return 1;
    }
    usleep(1000);
  }
  FILE *f = std::fopen(name.c_str(), "rb");
  if (!f) return 1;
  const size_t got = bytes ? std::fread(buf, 1, bytes, f) : 0;
  std::fclose(f);
  unlink(name.c_str());
  return got == bytes ? 0 : 1;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t c,
                           cudaStream_t s) {
  // every rank sends its contribution to every other rank and combines them in rank order (ncclSum = 0, ncclMax = 2)
  if (op != 0 && op != 2) return 1;
  const size_t bytes = count * type_size(type);
  std::vector<char> mine(bytes);
  std::memcpy(mine.data(), send, bytes);
  for (int r = 0; r < c->nranks; r++)
    if (r != c->rank && ncclSend(mine.data(), count, type, r, c, s)) return 1;
  std::vector<char> acc(bytes, 0), other(bytes);
  for (int r = 0; r < c->nranks; r++) {
    const char *src = mine.data();
    if (r != c->rank) {
      if (ncclRecv(other.data(), count, type, r, c, s)) return 1;
      src = other.data();
    }
    for (size_t i = 0; i < count; i++) {
      if (type == 8) {
        double &a = reinterpret_cast<double *>(acc.data())[i];
        const double v = reinterpret_cast<const double *>(src)[i];
        a = (r == 0) ? v : (op == 0 ? a + v : (v > a ? v : a));
      } else if (type == 2) {
        int &a = reinterpret_cast<int *>(acc.data())[i];
        const int v = reinterpret_cast<const int *>(src)[i];
        a = (r == 0) ? v : (op == 0 ? a + v : (v > a ? v : a));
      } else {
        return 1;
      }
    }
  }
  std::memcpy(recv, acc.data(), bytes);
  return 0;
}

const char *ncclGetErrorString(ncclResult_t r) { return r == 0 ? "no error" : "fake_nccl: failure"; }

}  // extern "C"
