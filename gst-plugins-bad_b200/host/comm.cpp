// comm.cpp - multi-GPU row-shard plumbing: one process per GPU, NCCL over NVLink.
//
// The reference has no multi-device path at all (only CUDA P2P enable in
// sys/nvcodec/gstcudacontext.c:223-247); row-block sharding with a halo exchange for the
// stencil elements is new. Point/LUT elements shard with no collective;
// bayer2rgb needs 1 halo row of its u8 input, dilate 1 row below, gaussianblur `center`
// rows (u8 input rows are exchanged and the horizontal pass recomputed, 4 B/px instead
// of 16 B/px of fp32 intermediates).
// NCCL is dlopen'd (libnccl.so.2) so that the single-GPU library has no hard dependency,
// the way the reference dlopens libcuda (sys/nvcodec/gstcudaloader.c:30-34).
#include "../csrc/common.cuh"
#include <dlfcn.h>
#include <string.h>
#include <mutex>
#include <string>

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclInt32 = 2, ncclSum = 0 };

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId) (ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank) (ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy) (ncclComm_t) = nullptr;
  ncclResult_t (*Send) (const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv) (void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart) () = nullptr;
  ncclResult_t (*GroupEnd) () = nullptr;
  ncclResult_t (*AllReduce) (const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString) (ncclResult_t) = nullptr;
  bool ok = false;
  std::string load_error;          // dlerror() text saved at load time (a second dlerror() call returns NULL)
};

NcclApi &nccl () {
  static NcclApi api;
  static std::once_flag once;
  std::call_once (once, [] () {
    const char *names[] = { "libnccl.so.2", "libnccl.so", nullptr };   // soname match reuses a copy torch already loaded
    for (int i = 0; names[i] && !api.handle; i++) api.handle = dlopen (names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) { const char *e = dlerror (); api.load_error = e ? e : "dlopen failed"; return; }
#define LOAD(field, sym) api.field = (decltype (api.field)) dlsym (api.handle, sym); if (!api.field) { api.load_error = std::string ("missing symbol ") + sym; return; }
    LOAD (GetUniqueId, "ncclGetUniqueId");
    LOAD (CommInitRank, "ncclCommInitRank");
    LOAD (CommDestroy, "ncclCommDestroy");
    LOAD (Send, "ncclSend");
    LOAD (Recv, "ncclRecv");
    LOAD (GroupStart, "ncclGroupStart");
    LOAD (GroupEnd, "ncclGroupEnd");
    LOAD (AllReduce, "ncclAllReduce");
    LOAD (GetErrorString, "ncclGetErrorString");
#undef LOAD
    api.ok = true;
  });
  return api;
}

#define NCCL_CHECK(expr)                                                              \
  do {                                                                                \
    ncclResult_t _r = (expr);                                                         \
    if (_r != 0) {                                                                    \
      b200vf_set_error ("%s failed: %s", #expr, nccl ().GetErrorString (_r));         \
      return B200VF_E_NCCL;                                                           \
    }                                                                                 \
  } while (0)

// inside ncclGroupStart .. ncclGroupEnd: close the group before returning, or later calls on this thread would queue
// into a group that never ends
#define NCCL_CHECK_IN_GROUP(expr)                                                     \
  do {                                                                                \
    ncclResult_t _r = (expr);                                                         \
    if (_r != 0) {                                                                    \
      b200vf_set_error ("%s failed: %s", #expr, nccl ().GetErrorString (_r));         \
      nccl ().GroupEnd ();                                                            \
      return B200VF_E_NCCL;                                                           \
    }                                                                                 \
  } while (0)

}  // namespace

struct b200vf_comm {
  b200vf_ctx *ctx = nullptr;
  int device = -1;                 // cached: destroy must not touch a context that was destroyed first
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int *d_flag = nullptr;
  uint8_t *scratch = nullptr;      // packed halos: [send_up | send_down | recv_up | recv_down]
  size_t scratch_bytes = 0;
  // split-phase exchange (halo_begin / halo_end): the exchange runs on the communicator's own stream
  cudaStream_t side = nullptr;
  cudaEvent_t ev_in = nullptr, ev_done = nullptr;
};

B200VF_API int b200vf_comm_unique_id (uint8_t id_out[128]) {
  B200VF_REQUIRE (id_out, B200VF_E_INVAL, "comm_unique_id: NULL argument");
  B200VF_REQUIRE (nccl ().ok, B200VF_E_NCCL, "libnccl.so.2 could not be loaded: %s", nccl ().load_error.c_str ());
  ncclUniqueId id;
  NCCL_CHECK (nccl ().GetUniqueId (&id));
  memcpy (id_out, id.internal, 128);
  return B200VF_OK;
}

B200VF_API int b200vf_comm_create (b200vf_ctx *ctx, const uint8_t id[128], int rank, int nranks, b200vf_comm **out) {
  B200VF_REQUIRE (ctx && id && out && nranks >= 1 && rank >= 0 && rank < nranks, B200VF_E_INVAL, "comm_create: bad argument");
  B200VF_REQUIRE (nccl ().ok, B200VF_E_NCCL, "libnccl.so.2 could not be loaded: %s", nccl ().load_error.c_str ());
  B200VF_CHECK_CUDA (cudaSetDevice (ctx->device));
  ncclUniqueId uid;
  memcpy (uid.internal, id, 128);
  b200vf_comm *c = new b200vf_comm ();
  c->ctx = ctx; c->device = ctx->device; c->rank = rank; c->nranks = nranks;
  ncclResult_t r = nccl ().CommInitRank (&c->comm, nranks, uid, rank);
  if (r != 0) {
    b200vf_set_error ("ncclCommInitRank failed: %s", nccl ().GetErrorString (r));
    delete c;
    return B200VF_E_NCCL;
  }
  if (b200vf_malloc (ctx, sizeof (int), (void **) &c->d_flag) != B200VF_OK) { nccl ().CommDestroy (c->comm); delete c; return B200VF_E_NOMEM; }
  *out = c;
  return B200VF_OK;
}

B200VF_API void b200vf_comm_destroy (b200vf_comm *comm) {
  if (!comm) return;
  cudaSetDevice (comm->device);
  if (comm->comm) nccl ().CommDestroy (comm->comm);
  if (comm->d_flag) cudaFree (comm->d_flag);
  if (comm->scratch) cudaFree (comm->scratch);
  if (comm->side) cudaStreamDestroy (comm->side);
  if (comm->ev_in) cudaEventDestroy (comm->ev_in);
  if (comm->ev_done) cudaEventDestroy (comm->ev_done);
  delete comm;
}

// contiguous row blocks, even row0 (Bayer 2x2 phase), remainder pairs to the first ranks
B200VF_API int b200vf_shard_rows (int height, int rank, int nranks, int *row0, int *rows) {
  B200VF_REQUIRE (row0 && rows && height > 0 && nranks >= 1 && rank >= 0 && rank < nranks, B200VF_E_INVAL, "shard_rows: bad argument");
  int pairs = height / 2, per = pairs / nranks, rem = pairs % nranks;
  B200VF_REQUIRE (per >= 2, B200VF_E_INVAL, "shard_rows: %d rows over %d ranks leaves shards under 4 rows", height, nranks);
  *row0 = 2 * (rank * per + (rank < rem ? rank : rem));
  *rows = 2 * (per + (rank < rem ? 1 : 0));
  if (rank == nranks - 1) *rows += height & 1;
  return B200VF_OK;
}

B200VF_API int b200vf_comm_halo_exchange (b200vf_comm *comm, uint8_t *d_buf, size_t row_bytes, int rows, int halo,
    size_t frame_stride, int nframes, void *stream)
{
  B200VF_REQUIRE (comm && d_buf && row_bytes > 0 && rows >= halo && halo >= 0 && nframes > 0, B200VF_E_INVAL, "halo_exchange: bad argument");
  B200VF_REQUIRE (frame_stride >= row_bytes * (size_t) (rows + 2 * halo) || nframes == 1, B200VF_E_INVAL, "halo_exchange: frame stride");
  if (halo == 0 || comm->nranks == 1) return B200VF_OK;
  cudaStream_t s = b200vf_stream (comm->ctx, stream);
  const int up = comm->rank - 1, down = comm->rank + 1;
  const size_t hb = row_bytes * halo;            // halo bytes of one frame
  const size_t part = hb * nframes;              // ... of the whole batch, packed
  // The halo rows of a batch are `nframes` strided pieces; they are packed into one contiguous
  // message per neighbour (strided D2D copy), so a step costs one ncclSend/ncclRecv pair per
  // neighbour however many frames are in flight, and unpacked into the head-room the same way.
  if (comm->scratch_bytes < 4 * part) {
    if (comm->scratch) B200VF_CHECK_CUDA (cudaFree (comm->scratch));
    comm->scratch = nullptr;
    comm->scratch_bytes = 0;
    cudaError_t e = cudaMalloc ((void **) &comm->scratch, 4 * part);
    B200VF_REQUIRE (e == cudaSuccess, B200VF_E_NOMEM, "halo_exchange: cudaMalloc(%zu): %s", 4 * part, cudaGetErrorString (e));
    comm->scratch_bytes = 4 * part;
  }
  uint8_t *send_up = comm->scratch, *send_down = send_up + part, *recv_up = send_down + part, *recv_down = recv_up + part;
  uint8_t *top_halo = d_buf, *first = d_buf + hb, *last = d_buf + row_bytes * rows, *bottom_halo = d_buf + hb + row_bytes * rows;
  if (up >= 0)
    B200VF_CHECK_CUDA (cudaMemcpy2DAsync (send_up, hb, first, frame_stride, hb, nframes, cudaMemcpyDeviceToDevice, s));
  if (down < comm->nranks)
    B200VF_CHECK_CUDA (cudaMemcpy2DAsync (send_down, hb, last, frame_stride, hb, nframes, cudaMemcpyDeviceToDevice, s));
  NCCL_CHECK (nccl ().GroupStart ());
  if (up >= 0) {
    NCCL_CHECK_IN_GROUP (nccl ().Send (send_up, part, ncclUint8, up, comm->comm, s));
    NCCL_CHECK_IN_GROUP (nccl ().Recv (recv_up, part, ncclUint8, up, comm->comm, s));
  }
  if (down < comm->nranks) {
    NCCL_CHECK_IN_GROUP (nccl ().Send (send_down, part, ncclUint8, down, comm->comm, s));
    NCCL_CHECK_IN_GROUP (nccl ().Recv (recv_down, part, ncclUint8, down, comm->comm, s));
  }
  NCCL_CHECK (nccl ().GroupEnd ());
  if (up >= 0)
    B200VF_CHECK_CUDA (cudaMemcpy2DAsync (top_halo, frame_stride, recv_up, hb, hb, nframes, cudaMemcpyDeviceToDevice, s));
  if (down < comm->nranks)
    B200VF_CHECK_CUDA (cudaMemcpy2DAsync (bottom_halo, frame_stride, recv_down, hb, hb, nframes, cudaMemcpyDeviceToDevice, s));
  return B200VF_OK;
}

// Split-phase halo exchange: pack -> ncclSend/ncclRecv -> unpack run on the communicator's own stream, ordered after
// everything queued on `stream` so far (the producers of the shard rows); halo_end makes `stream` wait for it. Between
// the two the caller launches the kernels that do not read the halo rows - every tile row but a shard's first and last -
// so that the exchange is off the critical path (it is 1 row in ~270 for bayer2rgb at 8 GPUs, but pack + NCCL + unpack
// cost ~0.2 ms of latency per step when serialised ahead of the kernel).
B200VF_API int b200vf_comm_halo_begin (b200vf_comm *comm, uint8_t *d_buf, size_t row_bytes, int rows, int halo,
    size_t frame_stride, int nframes, void *stream)
{
  B200VF_REQUIRE (comm, B200VF_E_INVAL, "halo_begin: NULL argument");
  cudaStream_t s = b200vf_stream (comm->ctx, stream);
  if (!comm->side) {
    B200VF_CHECK_CUDA (cudaStreamCreateWithFlags (&comm->side, cudaStreamNonBlocking));
    B200VF_CHECK_CUDA (cudaEventCreateWithFlags (&comm->ev_in, cudaEventDisableTiming));
    B200VF_CHECK_CUDA (cudaEventCreateWithFlags (&comm->ev_done, cudaEventDisableTiming));
  }
  B200VF_CHECK_CUDA (cudaEventRecord (comm->ev_in, s));
  B200VF_CHECK_CUDA (cudaStreamWaitEvent (comm->side, comm->ev_in, 0));
  int rc = b200vf_comm_halo_exchange (comm, d_buf, row_bytes, rows, halo, frame_stride, nframes, comm->side);
  if (rc) return rc;
  B200VF_CHECK_CUDA (cudaEventRecord (comm->ev_done, comm->side));
  return B200VF_OK;
}
B200VF_API int b200vf_comm_halo_end (b200vf_comm *comm, void *stream)
{
  B200VF_REQUIRE (comm && comm->ev_done, B200VF_E_INVAL, "halo_end: no exchange in flight");
  B200VF_CHECK_CUDA (cudaStreamWaitEvent (b200vf_stream (comm->ctx, stream), comm->ev_done, 0));
  return B200VF_OK;
}

B200VF_API int b200vf_comm_allgather_rows (b200vf_comm *comm, uint8_t *d_full, size_t row_bytes, int full_rows,
    size_t frame_stride, int nframes, void *stream)
{
  B200VF_REQUIRE (comm && d_full && row_bytes > 0 && full_rows > 0 && nframes > 0, B200VF_E_INVAL, "allgather_rows: bad argument");
  if (comm->nranks == 1) return B200VF_OK;
  cudaStream_t s = b200vf_stream (comm->ctx, stream);
  int my0 = 0, myn = 0;
  int rc = b200vf_shard_rows (full_rows, comm->rank, comm->nranks, &my0, &myn);
  if (rc) return rc;
  for (int f = 0; f < nframes; f++) {
    uint8_t *base = d_full + (size_t) f * frame_stride;
    NCCL_CHECK (nccl ().GroupStart ());
    for (int r = 0; r < comm->nranks; r++) {
      if (r == comm->rank) continue;
      int r0 = 0, rn = 0;
      rc = b200vf_shard_rows (full_rows, r, comm->nranks, &r0, &rn);
      if (rc) { nccl ().GroupEnd (); return rc; }
      NCCL_CHECK_IN_GROUP (nccl ().Send (base + (size_t) my0 * row_bytes, (size_t) myn * row_bytes, ncclUint8, r, comm->comm, s));
      NCCL_CHECK_IN_GROUP (nccl ().Recv (base + (size_t) r0 * row_bytes, (size_t) rn * row_bytes, ncclUint8, r, comm->comm, s));
    }
    NCCL_CHECK (nccl ().GroupEnd ());
  }
  return B200VF_OK;
}

B200VF_API int b200vf_comm_exchange_rows (b200vf_comm *comm, uint8_t *d_full, size_t row_bytes, int full_rows,
    const int *need_lo, const int *need_hi, size_t frame_stride, int nframes, void *stream)
{
  B200VF_REQUIRE (comm && d_full && need_lo && need_hi && row_bytes > 0 && full_rows > 0 && nframes > 0, B200VF_E_INVAL,
      "exchange_rows: bad argument");
  if (comm->nranks == 1) return B200VF_OK;
  cudaStream_t s = b200vf_stream (comm->ctx, stream);
  int my0 = 0, myn = 0;
  int rc = b200vf_shard_rows (full_rows, comm->rank, comm->nranks, &my0, &myn);
  if (rc) return rc;
  auto clip = [] (int a0, int a1, int b0, int b1, int &o0, int &o1) { o0 = a0 > b0 ? a0 : b0; o1 = a1 < b1 ? a1 : b1; return o1 > o0; };
  for (int f = 0; f < nframes; f++) {
    uint8_t *base = d_full + (size_t) f * frame_stride;
    NCCL_CHECK (nccl ().GroupStart ());
    for (int r = 0; r < comm->nranks; r++) {
      if (r == comm->rank) continue;
      int r0 = 0, rn = 0, a, b;
      rc = b200vf_shard_rows (full_rows, r, comm->nranks, &r0, &rn);
      if (rc) { nccl ().GroupEnd (); return rc; }
      if (clip (my0, my0 + myn, need_lo[r], need_hi[r], a, b))           // my rows that peer r reads
        NCCL_CHECK_IN_GROUP (nccl ().Send (base + (size_t) a * row_bytes, (size_t) (b - a) * row_bytes, ncclUint8, r, comm->comm, s));
      if (clip (r0, r0 + rn, need_lo[comm->rank], need_hi[comm->rank], a, b))   // peer r's rows that I read
        NCCL_CHECK_IN_GROUP (nccl ().Recv (base + (size_t) a * row_bytes, (size_t) (b - a) * row_bytes, ncclUint8, r, comm->comm, s));
    }
    NCCL_CHECK (nccl ().GroupEnd ());
  }
  return B200VF_OK;
}

B200VF_API int b200vf_comm_barrier (b200vf_comm *comm, void *stream) {
  B200VF_REQUIRE (comm, B200VF_E_INVAL, "comm_barrier: NULL argument");
  cudaStream_t s = b200vf_stream (comm->ctx, stream);
  NCCL_CHECK (nccl ().AllReduce (comm->d_flag, comm->d_flag, 1, ncclInt32, ncclSum, comm->comm, s));
  B200VF_CHECK_CUDA (cudaStreamSynchronize (s));
  return B200VF_OK;
}
