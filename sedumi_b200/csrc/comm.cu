// comm.cu -- the data-plane collective of the sharded hot path, inside the library.
//
// One process per GPU (torchrun, MPI, or N interpreter workers): rank 0 asks for a unique id, the host moves those
// 128 bytes to the other ranks by whatever means it has (bench.py: a torch.distributed broadcast; a MATLAB host: a
// file or parpool message), and every rank calls sb200_comm_init_rank.  From then on the collectives are ordinary
// asynchronous work on the library stream -- they are captured into the iteration's CUDA graph together with the
// kernels they separate, so a sharded iteration is still ONE graph launch.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the soname both the system package and the copy bundled with
// PyTorch carry): a single-GPU user of the MEX plugins never needs the library, and inside a PyTorch process the
// already loaded copy is reused, so that two NCCL instances never share a process.
#include <dlfcn.h>
#include "sb_internal.h"

namespace sb {
namespace {
typedef void *nccl_comm_t;
struct nccl_uid { char internal[128]; };
typedef int (*fn_get_version)(int *);
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(nccl_comm_t *, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char *(*fn_errstr)(int);
typedef int (*fn_group)(void);
enum { NCCL_F64 = 8, NCCL_SUM = 0 };     // ncclFloat64, ncclSum (nccl.h; stable since NCCL 2.0)

struct Nccl {
  void *h = nullptr;
  fn_get_version get_version = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  nccl_comm_t comm = nullptr;
  int nranks = 1, rank = 0, version = 0;
  int64_t calls = 0, bytes = 0;
};
Nccl g_nccl;

int nccl_load() {
  Nccl &N = g_nccl;
  if (N.h) return 0;
  const char *names[] = {getenv("SB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    N.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (N.h) break;
  }
  SB_CHECK(N.h, "multi-GPU: cannot load libnccl.so.2 (%s); set SB200_NCCL_LIB to its path", dlerror());
#define SB_SYM(field, name)                                                             \
  N.field = (decltype(N.field))dlsym(N.h, name);                                        \
  SB_CHECK(N.field, "multi-GPU: %s not found in the NCCL library", name)
  SB_SYM(get_version, "ncclGetVersion");
  SB_SYM(get_uid, "ncclGetUniqueId");
  SB_SYM(init_rank, "ncclCommInitRank");
  SB_SYM(allreduce, "ncclAllReduce");
  SB_SYM(destroy, "ncclCommDestroy");
  SB_SYM(errstr, "ncclGetErrorString");
  SB_SYM(group_start, "ncclGroupStart");
  SB_SYM(group_end, "ncclGroupEnd");
#undef SB_SYM
  N.get_version(&N.version);
  return 0;
}
#define SB_NCCL(call)                                                                  \
  do {                                                                                 \
    int _r = (call);                                                                   \
    if (_r != 0) { sb::set_error("%s failed: %s", #call, g_nccl.errstr ? g_nccl.errstr(_r) : "?"); return 1; } \
  } while (0)
}  // namespace
}  // namespace sb

using namespace sb;

extern "C" {

// 128 bytes identifying a new communicator; call on ONE rank and hand the bytes to all ranks.
int sb200_comm_unique_id(void *id128) {
  SB_TRY(nccl_load());
  nccl_uid id;
  SB_NCCL(g_nccl.get_uid(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

// Join the communicator as `rank` of `nranks` with the library's device (sb200_init) and stream.  Collective.
int sb200_comm_init_rank(int nranks, int rank, const void *id128) {
  SB_TRY(ensure_init());
  SB_TRY(nccl_load());
  SB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "comm_init_rank: rank %d of %d", rank, nranks);
  if (g_nccl.comm) { g_nccl.destroy(g_nccl.comm); g_nccl.comm = nullptr; }
  nccl_uid id;
  memcpy(&id, id128, sizeof id);
  SB_CUDA(cudaSetDevice(ctx().device));
  SB_NCCL(g_nccl.init_rank(&g_nccl.comm, nranks, id, rank));
  g_nccl.nranks = nranks; g_nccl.rank = rank;
  return 0;
}
int sb200_comm_size(void) { return g_nccl.comm ? g_nccl.nranks : 1; }
int sb200_comm_rank(void) { return g_nccl.comm ? g_nccl.rank : 0; }
int sb200_comm_nccl_version(void) { return g_nccl.version; }
// collectives issued / bytes reduced by this rank so far (bench evidence that the sharded path really communicates)
int sb200_comm_stats(int64_t *calls, int64_t *bytes) {
  if (calls) *calls = g_nccl.calls;
  if (bytes) *bytes = g_nccl.bytes;
  return 0;
}

// In-place sum over all ranks of `count` doubles at buf_dev, enqueued on the library stream (capturable).
// A world of one is the identity and launches nothing.
int sb200_allreduce_sum_dev(double *buf_dev, int64_t count) {
  SB_TRY(ensure_init());
  if (count <= 0) return 0;
  if (!g_nccl.comm || g_nccl.nranks == 1) return 0;
  SB_NCCL(g_nccl.allreduce(buf_dev, buf_dev, (size_t)count, NCCL_F64, NCCL_SUM, g_nccl.comm, ctx().stream));
  g_nccl.calls++; g_nccl.bytes += 8 * count;
  if (ctx().profiling) prof_mark("nccl_allreduce");
  return 0;
}
// Two buffers reduced as one group (one fused launch): ADA values + absd at the Schur-assembly boundary.
int sb200_allreduce_sum2_dev(double *a_dev, int64_t na, double *b_dev, int64_t nb) {
  SB_TRY(ensure_init());
  if (!g_nccl.comm || g_nccl.nranks == 1) return 0;
  SB_NCCL(g_nccl.group_start());
  if (na > 0) SB_NCCL(g_nccl.allreduce(a_dev, a_dev, (size_t)na, NCCL_F64, NCCL_SUM, g_nccl.comm, ctx().stream));
  if (nb > 0) SB_NCCL(g_nccl.allreduce(b_dev, b_dev, (size_t)nb, NCCL_F64, NCCL_SUM, g_nccl.comm, ctx().stream));
  SB_NCCL(g_nccl.group_end());
  g_nccl.calls++; g_nccl.bytes += 8 * (na + nb);
  if (ctx().profiling) prof_mark("nccl_allreduce");
  return 0;
}

int sb200_comm_destroy(void) {
  if (g_nccl.comm) {
    if (ctx().stream) cudaStreamSynchronize(ctx().stream);
    g_nccl.destroy(g_nccl.comm);
    g_nccl.comm = nullptr;
  }
  g_nccl.nranks = 1; g_nccl.rank = 0;
  return 0;
}

}  // extern "C"
