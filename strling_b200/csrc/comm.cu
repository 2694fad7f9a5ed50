// Multi-GPU entry points of libstrgpu.so (include/strgpu.h, "sharded clustering"): one process per GPU, one NCCL
// communicator per context, NVLink / NVSwitch underneath.
//
// The reference has no parallelism of its own beyond running `strling merge --chromosome C` once per chromosome
// (merge.nim:52,89; pipelines/strling-joint.groovy:7-12), which works because reads are grouped by (tid, repeat) before
// anything else happens to them (call.nim:124-125, merge.nim:125) and buckets never interact.  The same fact shards the
// cluster stage exactly: every bucket has an owner rank,
//   1. each rank partitions its STR-read records by owner -- one stable counting pass on the owner "digit" (histogram, scan,
//      tile-local scatter: the same scheme as the radix sort passes) straight into the NCCL send buffer,
//   2. the records travel to their owners (grouped ncclSend / ncclRecv of fixed-capacity slots; the per-pair counts travel in
//      an all-gather and are consumed ON THE DEVICE, so the host never waits for a count),
//   3. each rank clusters what it owns (run_cluster with a device-side record count; rank-major arrival order ==
//      concatenation order, so ties sort exactly as on one GPU over the concatenated input),
//   4. the 48-byte cluster records are all-gathered (the only collective the north-star names) and put into single-GPU order
//      by one stable sort on (tid, unit) -- a bucket lives on one rank, so its clusters are already in position order.
// Nothing in steps 1-4 synchronises with the host.  NCCL is loaded with dlopen at strgpu_comm_init, so single-GPU users of
// the library (the CLI, the scan) do not need it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.cuh"

namespace strgpu_internal {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    // inside a process that already loaded NCCL (torch) this returns that copy; otherwise the system library
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
#define LOAD(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym))
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllGather, "ncclAllGather");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.Send || !api.Recv || !api.GroupStart ||
        !api.GroupEnd || !api.GetErrorString) {
      dlclose(api.lib);
      api.lib = nullptr;
    }
  });
  return api.lib ? &api : nullptr;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  DevBuf send, recv, own, counts, gath, cat, small, peer_tab;
  // peer-to-peer exchange: every rank's receive buffer is mapped into every other rank (cudaIpc over NVLink / NVSwitch), so the
  // partition kernel stores each record straight into its owner's memory -- partition and exchange are ONE kernel, only the
  // records themselves cross the links (no padded slots), and the count all-gather that follows is the barrier
  int tab_flip = 0;
  void *h_tab = nullptr;     // pinned copy of the destination table for captured graphs
  GraphSlot graph;
  void *peer_recv[256] = {nullptr};   // [rank] -> that rank's recv buffer in this process's address space (own entry: recv.p)
  bool p2p = false;
  // small (uint32 words): [0 .. W) send counts, [W .. W + W*W) count matrix [src][dst], then: records owned, local bounds,
  // W gathered bound counts, total bounds, overflow flags (bit 0 exchange slot, bit 1 gather slot)
};

#define NC(ctx, call)                                                                                                         \
  do {                                                                                                                        \
    ncclResult_t r_ = (call);                                                                                                 \
    if (r_ != ncclSuccess) return fail(ctx, STRGPU_ERR_CUDA, "%s: %s", #call, nccl_api()->GetErrorString(r_));                 \
  } while (0)

void comm_release(strgpu_ctx *ctx) {
  if (!ctx || !ctx->comm) return;
  Comm *c = ctx->comm;
  cudaDeviceSynchronize();
  if (c->graph.exec) cudaGraphExecDestroy(c->graph.exec);
  if (c->h_tab) cudaFreeHost(c->h_tab);
  for (int r = 0; r < c->world; r++)
    if (c->p2p && r != c->rank && c->peer_recv[r]) cudaIpcCloseMemHandle(c->peer_recv[r]);
  if (c->comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
  for (DevBuf *b : {&c->send, &c->recv, &c->own, &c->counts, &c->gath, &c->cat, &c->small, &c->peer_tab})
    if (b->p) cudaFree(b->p);
  delete c;
  ctx->comm = nullptr;
}

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kPartWarps = 8;
constexpr uint32_t kPartTile = 2048;                      // treads per CTA
constexpr uint32_t kPartSpan = kPartTile / kPartWarps;    // per warp
constexpr int kMaxWorld = 256;

// owner rank of a (tid, repeat) bucket: any deterministic function of the bucket key shards the stage exactly
__device__ __forceinline__ uint32_t bucket_owner(const strgpu_tread &t, uint32_t world) {
  unsigned long long k = (unsigned long long)(uint32_t)t.tid;
#pragma unroll
  for (int j = 0; j < 6; j++) k = k * 0x100000001b3ull ^ (unsigned long long)(unsigned char)t.repeat[j];
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)(k % world);
}

__global__ void __launch_bounds__(kPartWarps * 32) part_histogram(const strgpu_tread *__restrict__ treads, uint32_t n, uint32_t world,
                                                                  uint32_t n_tiles, uint32_t *__restrict__ counts) {
  __shared__ uint32_t hist[kMaxWorld];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t beg = blockIdx.x * kPartTile, end = min(n, beg + kPartTile);
  for (uint32_t i = beg + threadIdx.x; i < end; i += kPartWarps * 32) atomicAdd(&hist[bucket_owner(treads[i], world)], 1u);
  __syncthreads();
  if (threadIdx.x < world) counts[(size_t)threadIdx.x * n_tiles + blockIdx.x] = hist[threadIdx.x];
}

// offsets = exclusive scan of counts in (owner, tile) order.  Slot of a record = its rank among the records this shard has for
// that owner (input order: stable).  The tile is first sorted by owner in shared memory, then copied out run by run, so the
// stores to one owner are consecutive 24-byte records: dst_tab[owner] + slot.  dst_tab points either into the local send
// buffer (NCCL exchange) or -- peer-to-peer -- straight into the owner's receive buffer on another GPU.
// sendcnt[owner] is written by the last tile.
constexpr int kPartSmem = kPartTile * sizeof(strgpu_tread) + kPartTile + kPartWarps * kMaxWorld * 4 + 2 * kMaxWorld * 4;

__global__ void __launch_bounds__(kPartWarps * 32) part_scatter(const strgpu_tread *__restrict__ treads, uint32_t n, uint32_t world,
                                                                uint32_t n_tiles, const uint32_t *__restrict__ offsets, uint32_t pair_cap,
                                                                strgpu_tread *const *__restrict__ dst_tab, uint32_t *__restrict__ sendcnt,
                                                                uint32_t *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char part_smem[];
  unsigned long long *stage = reinterpret_cast<unsigned long long *>(part_smem);                     // kPartTile records, 3 words each
  unsigned char *own_of = part_smem + kPartTile * sizeof(strgpu_tread);                               // owner of the record at a tile position
  uint32_t(*whist)[kMaxWorld] = reinterpret_cast<uint32_t(*)[kMaxWorld]>(own_of + kPartTile);        // [warp][owner]
  uint32_t *lbase = reinterpret_cast<uint32_t *>(whist + kPartWarps);                                 // tile-local first position of an owner
  uint32_t *gbase = lbase + kMaxWorld;                                                                // its first slot in the owner's segment
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t beg = blockIdx.x * kPartTile, end = min(n, beg + kPartTile);
  const uint32_t wbeg = min(end, beg + warp * kPartSpan), wend = min(end, wbeg + kPartSpan);
  for (int w = 0; w < kPartWarps; w++) whist[w][tid] = 0;
  __syncthreads();
  for (uint32_t i = wbeg + lane; i < wend; i += 32) atomicAdd(&whist[warp][bucket_owner(treads[i], world)], 1u);
  __syncthreads();
  {
    uint32_t run = 0;
    for (int w = 0; w < kPartWarps; w++) {
      const uint32_t c = whist[w][tid];
      whist[w][tid] = run;
      run += c;
    }
    // exclusive scan of the owners' tile totals (256 threads: one warp-shuffle scan + warp totals through lbase)
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) gbase[warp] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; w++) before += gbase[w];
    __syncthreads();
    const uint32_t ex = before + inc - run;
    lbase[tid] = ex;
    for (int w = 0; w < kPartWarps; w++) whist[w][tid] += ex;
    uint32_t first = 0;
    if ((uint32_t)tid < world) {
      first = offsets[(size_t)tid * n_tiles + blockIdx.x] - offsets[(size_t)tid * n_tiles];
      if (blockIdx.x == n_tiles - 1) {
        sendcnt[tid] = first + run;   // records for this owner in the whole shard
        if (first + run > pair_cap) atomicOr(flags, 1u);
      }
    }
    gbase[tid] = first;
  }
  __syncthreads();
  const uint32_t lane_lt = (1u << lane) - 1u;
  for (uint32_t base = wbeg; base < wend; base += 32) {
    const uint32_t i = base + lane;
    const bool valid = i < wend;
    unsigned long long w0 = 0, w1 = 0, w2 = 0;
    uint32_t o = 0x10000u + (uint32_t)lane;
    if (valid) {
      const unsigned long long *src = reinterpret_cast<const unsigned long long *>(treads + i);
      w0 = src[0]; w1 = src[1]; w2 = src[2];
      strgpu_tread t;
      reinterpret_cast<unsigned long long *>(&t)[0] = w0;
      reinterpret_cast<unsigned long long *>(&t)[1] = w1;
      reinterpret_cast<unsigned long long *>(&t)[2] = w2;
      o = bucket_owner(t, world);
    }
    const uint32_t grp = __match_any_sync(kFull, o);
    uint32_t local = 0;
    if (valid) local = whist[warp][o] + __popc(grp & lane_lt);   // lane order == input order: stable
    __syncwarp();
    if (valid && lane == 31 - __clz(grp)) whist[warp][o] += __popc(grp);
    __syncwarp();
    if (valid) {
      stage[3 * local] = w0; stage[3 * local + 1] = w1; stage[3 * local + 2] = w2;
      own_of[local] = (unsigned char)o;
    }
  }
  __syncthreads();
  // copy-out, 8 bytes per thread and consecutive threads on consecutive words: inside an owner's run the stores of a warp are
  // one contiguous 256-byte span (what NVLink wants when the destination is a peer's memory)
  for (uint32_t w = tid; w < 3u * (end - beg); w += kPartWarps * 32) {
    const uint32_t j = w / 3u, part = w - 3u * j;
    const uint32_t o = own_of[j];
    const uint32_t slot = gbase[o] + (j - lbase[o]);
    if (slot < pair_cap) reinterpret_cast<unsigned long long *>(dst_tab[o] + slot)[part] = stage[w];
  }
}

// cntmat[src][dst]: what src sends to dst.  The records this rank received, in rank-major order (== the order of the
// concatenated shards), become one contiguous array; *n_own = how many.
__global__ void compact_received(const strgpu_tread *__restrict__ recv, const uint32_t *__restrict__ cntmat, uint32_t world, uint32_t me,
                                 uint32_t pair_cap, strgpu_tread *__restrict__ own, uint32_t *__restrict__ n_own) {
  const uint32_t src = blockIdx.y;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t before = 0, total = 0;
  for (uint32_t j = 0; j < world; j++) {
    const uint32_t c = min(cntmat[j * world + me], pair_cap);
    if (j < src) before += c;
    total += c;
  }
  if (src == 0 && k == 0) *n_own = total;
  const uint32_t mine = min(cntmat[src * world + me], pair_cap);
  if (k < mine) own[before + k] = recv[(size_t)src * pair_cap + k];
}

// gathered[rank][rank_cap] bounds + per-rank counts -> contiguous rank-major array, *n_total
__global__ void compact_gathered(const strgpu_bounds *__restrict__ gath, const uint32_t *__restrict__ counts, uint32_t world, uint32_t rank_cap,
                                 strgpu_bounds *__restrict__ cat, uint32_t *__restrict__ n_total, uint32_t *__restrict__ flags) {
  const uint32_t src = blockIdx.y;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t before = 0, total = 0;
  bool over = false;
  for (uint32_t j = 0; j < world; j++) {
    over = over || counts[j] > rank_cap;
    const uint32_t c = min(counts[j], rank_cap);
    if (j < src) before += c;
    total += c;
  }
  if (src == 0 && k == 0) {
    *n_total = total;
    if (over) atomicOr(flags, 2u);
  }
  if (k < min(counts[src], rank_cap)) cat[before + k] = gath[(size_t)src * rank_cap + k];
}

}  // namespace
}  // namespace strgpu_internal

using namespace strgpu_internal;

extern "C" {

int strgpu_comm_unique_id(void *id_out) {
  if (!id_out) return STRGPU_ERR_INVALID;
  NcclApi *api = nccl_api();
  if (!api) return STRGPU_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == STRGPU_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return STRGPU_ERR_CUDA;
  std::memcpy(id_out, &id, sizeof(id));
  return STRGPU_OK;
}

int strgpu_comm_init(strgpu_ctx *ctx, int rank, int world, const void *id) {
  if (!ctx || !id || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(ctx, STRGPU_ERR_INVALID, "comm_init: rank %d of %d", rank, world);
  NcclApi *api = nccl_api();
  if (!api) return fail(ctx, STRGPU_ERR_CUDA, "comm_init: libnccl.so.2 could not be loaded");
  if (ctx->comm) comm_release(ctx);
  CU(ctx, cudaSetDevice(ctx->device));
  Comm *c = new (std::nothrow) Comm();
  if (!c) return fail(ctx, STRGPU_ERR_INVALID, "out of host memory");
  c->rank = rank;
  c->world = world;
  ncclUniqueId nid;
  std::memcpy(&nid, id, sizeof(nid));
  ncclResult_t r = api->CommInitRank(&c->comm, world, nid, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(ctx, STRGPU_ERR_CUDA, "ncclCommInitRank: %s", api->GetErrorString(r));
  }
  ctx->comm = c;
  return STRGPU_OK;
}

int strgpu_comm_info(const strgpu_ctx *ctx, int *rank, int *world) {
  if (!ctx || !ctx->comm) return STRGPU_ERR_INVALID;
  if (rank) *rank = ctx->comm->rank;
  if (world) *world = ctx->comm->world;
  return STRGPU_OK;
}

void strgpu_comm_destroy(strgpu_ctx *ctx) { comm_release(ctx); }

namespace strgpu_internal {
namespace {
// (Re)allocates the receive buffer and maps every rank's buffer into every other rank.  Collective and synchronising: it runs
// when a call needs a larger buffer than the last one, which every rank decides identically from the call's arguments.
int setup_recv(strgpu_ctx *ctx, Comm *c, size_t bytes, cudaStream_t st) {
  NcclApi *api = nccl_api();
  const int W = c->world;
  CU(ctx, cudaStreamSynchronize(st));
  for (int r = 0; r < W; r++)
    if (c->p2p && r != c->rank && c->peer_recv[r]) cudaIpcCloseMemHandle(c->peer_recv[r]);
  for (int r = 0; r < W; r++) c->peer_recv[r] = nullptr;
  c->p2p = false;
  // nobody may free a buffer a peer still has mapped: a collective round trip first
  int rc;
  if ((rc = ensure(ctx, c->small, 4096))) return rc;
  NC(ctx, api->AllGather(c->small.p, (char *)c->small.p + 1024, 4, ncclUint8, c->comm, st));
  CU(ctx, cudaStreamSynchronize(st));
  if ((rc = ensure(ctx, c->recv, bytes))) return rc;
  if ((rc = ensure(ctx, c->peer_tab, (size_t)2 * kMaxWorld * sizeof(void *)))) return rc;
  static const bool no_p2p = getenv("STRGPU_NO_P2P") != nullptr;
  // exchange the IPC handles (64 bytes each) through the communicator
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  static_assert(sizeof(Msg) == 80, "handle message");
  Msg mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.ok = (!no_p2p && cudaIpcGetMemHandle(&mine.h, c->recv.p) == cudaSuccess) ? 1 : 0;
  if (!mine.ok) cudaGetLastError();
  char *dbuf = nullptr;
  CU(ctx, cudaMalloc(&dbuf, sizeof(Msg) * (size_t)(W + 1)));
  CU(ctx, cudaMemcpyAsync(dbuf, &mine, sizeof(Msg), cudaMemcpyHostToDevice, st));
  NC(ctx, api->AllGather(dbuf, dbuf + sizeof(Msg), sizeof(Msg), ncclUint8, c->comm, st));
  std::vector<Msg> all((size_t)W);
  CU(ctx, cudaMemcpyAsync(all.data(), dbuf + sizeof(Msg), sizeof(Msg) * (size_t)W, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  bool ok = true;
  for (int r = 0; r < W; r++) ok = ok && all[(size_t)r].ok;
  int opened = 1;
  if (ok) {
    for (int r = 0; r < W && opened; r++) {
      if (r == c->rank) { c->peer_recv[r] = c->recv.p; continue; }
      if (cudaIpcOpenMemHandle(&c->peer_recv[r], all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        c->peer_recv[r] = nullptr;
        opened = 0;
      }
    }
  } else {
    opened = 0;
  }
  // every rank must have every mapping, or all fall back to the NCCL exchange
  CU(ctx, cudaMemcpyAsync(dbuf, &opened, 4, cudaMemcpyHostToDevice, st));
  NC(ctx, api->AllGather(dbuf, dbuf + sizeof(Msg), 4, ncclUint8, c->comm, st));
  std::vector<int> flags((size_t)W);
  CU(ctx, cudaMemcpyAsync(flags.data(), dbuf + sizeof(Msg), 4 * (size_t)W, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  cudaFree(dbuf);
  bool all_open = true;
  for (int r = 0; r < W; r++) all_open = all_open && flags[(size_t)r] == 1;
  if (!all_open) {
    for (int r = 0; r < W; r++) {
      if (r != c->rank && c->peer_recv[r]) cudaIpcCloseMemHandle(c->peer_recv[r]);
      c->peer_recv[r] = nullptr;
    }
  }
  c->p2p = all_open;
  return STRGPU_OK;
}
}  // namespace
}  // namespace strgpu_internal

static int sharded_enqueue(strgpu_ctx *ctx, const void *d_treads, uint32_t n, uint32_t max_n, uint32_t pair_capacity,
                           const strgpu_cluster_params *params, void *d_out, uint32_t cap, void *d_n_out, cudaStream_t st,
                           uint64_t *launches_out, bool capturing) {
  Comm *c = ctx->comm;
  NcclApi *api = nccl_api();
  const uint32_t W = (uint32_t)c->world, me = (uint32_t)c->rank;
  // slot per (source, destination) pair: a hash partition gives every owner ~ n / W of a shard; 25 % + 1024 of slack
  uint64_t pair_cap64 = pair_capacity ? pair_capacity : (uint64_t)max_n / W + (uint64_t)max_n / (4 * W) + 1024;
  if (pair_cap64 > max_n) pair_cap64 = max_n;
  if (pair_cap64 == 0) pair_cap64 = 1;
  const uint32_t pair_cap = (uint32_t)pair_cap64;
  const uint64_t own_max64 = (uint64_t)pair_cap * W;
  if (own_max64 >= (1ull << 29)) return fail(ctx, STRGPU_ERR_INVALID, "cluster_sharded: %llu records per rank exceed 2^29", (unsigned long long)own_max64);
  const uint32_t own_max = (uint32_t)own_max64;
  const uint32_t rank_cap = std::max<uint32_t>(1u, cap / W);   // bounds a rank may contribute to the gather
  const uint32_t n_tiles = std::max<uint32_t>(1u, (n + kPartTile - 1) / kPartTile);
  int rc;
  if ((size_t)own_max * sizeof(strgpu_tread) > c->recv.cap) {
    if (capturing) return fail(ctx, STRGPU_ERR_CUDA, "cluster_sharded: buffer growth during graph capture");
    if ((rc = setup_recv(ctx, c, (size_t)own_max * sizeof(strgpu_tread), st))) return rc;
    ctx->cluster_ws.gen++;   // captured graphs hold the old receive buffers
  }
  if (!c->p2p && (rc = ensure(ctx, c->send, (size_t)own_max * sizeof(strgpu_tread)))) return rc;
  if ((rc = ensure(ctx, c->own, (size_t)own_max * sizeof(strgpu_tread) + 64))) return rc;
  if ((rc = ensure(ctx, c->counts, (size_t)W * n_tiles * 4 + 64))) return rc;
  if ((rc = ensure(ctx, c->gath, (size_t)W * rank_cap * sizeof(strgpu_bounds)))) return rc;
  if ((rc = ensure(ctx, c->cat, (size_t)(W + 1) * rank_cap * sizeof(strgpu_bounds)))) return rc;
  const size_t small_words = (size_t)W + (size_t)W * W + W + 8;
  if ((rc = ensure(ctx, c->small, small_words * 4))) return rc;
  uint32_t *small = (uint32_t *)c->small.p;
  uint32_t *sendcnt = small, *cntmat = small + W, *n_own = cntmat + W * W, *n_loc = n_own + 1, *gcnt = n_loc + 1, *n_total = gcnt + W,
           *flags = n_total + 1;
  CU(ctx, cudaMemsetAsync(small, 0, small_words * 4, st));
  strgpu_tread *send = (strgpu_tread *)c->send.p, *recv = (strgpu_tread *)c->recv.p, *own = (strgpu_tread *)c->own.p;
  uint64_t launches = 0;
  // STRGPU_COMM_TIMING=1 (profiling only): per-phase device times on stderr; the call then synchronises
  static const bool timing_env = getenv("STRGPU_COMM_TIMING") != nullptr;
  const bool timing = timing_env && !capturing;
  cudaEvent_t ev[6] = {nullptr};
  auto mark = [&](int i) {
    if (!timing) return;
    cudaEventCreate(&ev[i]);
    cudaEventRecord(ev[i], st);
  };
  mark(0);

  // 1. stable partition by owner: into the owners' receive buffers (peer-to-peer stores over NVLink) or the local send slots
  {
    strgpu_tread *tab[kMaxWorld];
    for (uint32_t r = 0; r < W; r++)
      tab[r] = c->p2p ? (strgpu_tread *)c->peer_recv[r] + (size_t)me * pair_cap : send + (size_t)r * pair_cap;
    static std::once_flag attr_once[64];
    std::call_once(attr_once[ctx->device & 63], []() { cudaFuncSetAttribute(part_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, kPartSmem); });
    if ((rc = ensure(ctx, c->peer_tab, (size_t)2 * kMaxWorld * sizeof(void *)))) return rc;
    // two alternating halves so that a table still in use by the previous call's kernel is never overwritten
    c->tab_flip ^= 1;
    strgpu_tread **d_tab = (strgpu_tread **)c->peer_tab.p + (size_t)c->tab_flip * kMaxWorld;
    if (capturing) {
      // a captured copy re-reads its source at every replay: a pinned table that lives as long as the graph
      if (!c->h_tab && cudaMallocHost(&c->h_tab, kMaxWorld * sizeof(void *)) != cudaSuccess) return fail(ctx, STRGPU_ERR_CUDA, "cudaMallocHost");
      std::memcpy(c->h_tab, tab, W * sizeof(void *));
      CU(ctx, cudaMemcpyAsync(d_tab, c->h_tab, W * sizeof(void *), cudaMemcpyHostToDevice, st));
    } else {
      CU(ctx, cudaMemcpyAsync(d_tab, tab, W * sizeof(void *), cudaMemcpyHostToDevice, st));
    }
    uint32_t *counts = (uint32_t *)c->counts.p;
    part_histogram<<<n_tiles, kPartWarps * 32, 0, st>>>((const strgpu_tread *)d_treads, n, W, n_tiles, counts);
    CU(ctx, strgpu::scan_u32(ctx->cluster_ws, counts, counts, W * n_tiles, nullptr, st, &launches));
    part_scatter<<<n_tiles, kPartWarps * 32, kPartSmem, st>>>((const strgpu_tread *)d_treads, n, W, n_tiles, counts, pair_cap, d_tab, sendcnt, flags);
    launches += 2;
    CU(ctx, cudaGetLastError());
  }

  mark(1);
  // 2. counts (all-gather of the W send counts -> [src][dst] matrix) and records (fixed-capacity slots) to their owners
  // (peer-to-peer: the records are already in place; a rank's counts arrive after its partition kernel has finished, so this
  // all-gather is also the barrier that makes every peer's stores into recv visible)
  NC(ctx, api->AllGather(sendcnt, cntmat, W, ncclUint32, c->comm, st));
  if (!c->p2p) {
    NC(ctx, api->GroupStart());
    for (uint32_t peer = 0; peer < W; peer++) {
      if (peer == me) continue;
      NC(ctx, api->Send(send + (size_t)peer * pair_cap, (size_t)pair_cap * sizeof(strgpu_tread), ncclUint8, (int)peer, c->comm, st));
      NC(ctx, api->Recv(recv + (size_t)peer * pair_cap, (size_t)pair_cap * sizeof(strgpu_tread), ncclUint8, (int)peer, c->comm, st));
    }
    NC(ctx, api->GroupEnd());
    CU(ctx, cudaMemcpyAsync(recv + (size_t)me * pair_cap, send + (size_t)me * pair_cap, (size_t)pair_cap * sizeof(strgpu_tread),
                            cudaMemcpyDeviceToDevice, st));
  }
  {
    dim3 grid((pair_cap + 255) / 256, W);
    compact_received<<<grid, 256, 0, st>>>(recv, cntmat, W, me, pair_cap, own, n_own);
    launches++;
  }

  mark(2);
  // 3. cluster what this rank owns (record count known to the device only)
  strgpu_bounds *local = (strgpu_bounds *)c->cat.p + (size_t)W * rank_cap;   // the slot after the concatenation buffer
  CU(ctx, strgpu::run_cluster(ctx->cluster_ws, own, own_max, *params, local, rank_cap, n_loc, st, &launches, nullptr, n_own));

  mark(3);
  // 4. all-gather of the cluster records, then single-GPU order
  NC(ctx, api->AllGather(n_loc, gcnt, 1, ncclUint32, c->comm, st));
  NC(ctx, api->AllGather(local, c->gath.p, (size_t)rank_cap * sizeof(strgpu_bounds), ncclUint8, c->comm, st));
  {
    dim3 grid((rank_cap + 255) / 256, W);
    compact_gathered<<<grid, 256, 0, st>>>((const strgpu_bounds *)c->gath.p, gcnt, W, rank_cap, (strgpu_bounds *)c->cat.p, n_total, flags);
    launches++;
  }
  mark(4);
  CU(ctx, strgpu::sort_bounds_device(ctx->cluster_ws, (const strgpu_bounds *)c->cat.p, W * rank_cap, n_total, (strgpu_bounds *)d_out, cap,
                                     (uint32_t *)d_n_out, st, &launches));
  mark(5);
  if (timing) {
    cudaStreamSynchronize(st);
    float t[5];
    for (int i = 0; i < 5; i++) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
    fprintf(stderr, "[strgpu rank %u] sharded cluster: partition %.3f ms, exchange %.3f, cluster %.3f, gather %.3f, order %.3f (pair_cap %u, own_max %u, rank_cap %u, %s)\n",
            me, t[0], t[1], t[2], t[3], t[4], pair_cap, own_max, rank_cap, c->p2p ? "peer-to-peer stores" : "NCCL send/recv");
    for (auto &e : ev) cudaEventDestroy(e);
  }
  *launches_out += launches;
  return STRGPU_OK;
}

int strgpu_cluster_sharded_device(strgpu_ctx *ctx, const void *d_treads, uint32_t n, uint32_t max_n, uint32_t pair_capacity,
                                  const strgpu_cluster_params *params, void *d_out, uint32_t cap, void *d_n_out, void *cuda_stream) {
  if (!ctx || !params || !d_n_out || (n && !d_treads) || (cap && !d_out)) return fail(ctx, STRGPU_ERR_INVALID, "cluster_sharded: null argument");
  if (!ctx->comm) return fail(ctx, STRGPU_ERR_INVALID, "cluster_sharded: strgpu_comm_init has not been called");
  if (n > max_n) return fail(ctx, STRGPU_ERR_INVALID, "cluster_sharded: n %u > max_n %u", n, max_n);
  CU(ctx, cudaSetDevice(ctx->device));
  const strgpu_cluster_params p = *params;
  // Repeated calls with unchanged arguments are replayed as one CUDA graph (kernels, the small table upload and the NCCL
  // all-gathers; see GraphSlot).  Every rank of a job that repeats its call sees the same direct / capture / replay sequence.
  // STRGPU_SHARDED_GRAPH=0 keeps the sharded path on direct launches.
  static const bool graph_ok = !(getenv("STRGPU_SHARDED_GRAPH") && atoi(getenv("STRGPU_SHARDED_GRAPH")) == 0) && !getenv("STRGPU_COMM_TIMING");
  uint64_t key = hash_bytes(&d_treads, sizeof(d_treads));
  key = hash_bytes(&n, sizeof(n), key);
  key = hash_bytes(&max_n, sizeof(max_n), key);
  key = hash_bytes(&pair_capacity, sizeof(pair_capacity), key);
  key = hash_bytes(&p, sizeof(p), key);
  key = hash_bytes(&d_out, sizeof(d_out), key);
  key = hash_bytes(&cap, sizeof(cap), key);
  key = hash_bytes(&d_n_out, sizeof(d_n_out), key);
  cudaStream_t user = (cudaStream_t)cuda_stream;
  if (!graph_ok) {
    uint64_t l = 0;
    const int rc = sharded_enqueue(ctx, d_treads, n, max_n, pair_capacity, &p, d_out, cap, d_n_out, user, &l, false);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->launches += l;
    return rc;
  }
  return run_graphed(ctx, ctx->comm->graph, key, user, [&](cudaStream_t st, uint64_t *launches, bool capturing) -> int {
    return sharded_enqueue(ctx, d_treads, n, max_n, pair_capacity, &p, d_out, cap, d_n_out, st, launches, capturing);
  });
}

int strgpu_comm_status(strgpu_ctx *ctx, void *cuda_stream) {
  if (!ctx || !ctx->comm) return STRGPU_ERR_INVALID;
  Comm *c = ctx->comm;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize((cudaStream_t)cuda_stream));
  if (!c->small.p) return STRGPU_OK;
  const uint32_t W = (uint32_t)c->world;
  uint32_t flags = 0;
  CU(ctx, cudaMemcpy(&flags, (uint32_t *)c->small.p + W + W * W + 2 + W + 1, 4, cudaMemcpyDeviceToHost));
  if (flags & 1u) return fail(ctx, STRGPU_ERR_OVERFLOW, "cluster_sharded: a (source, owner) pair exceeded its exchange slot: pass a larger pair_capacity (max_n always fits)");
  if (flags & 2u) return fail(ctx, STRGPU_ERR_OVERFLOW, "cluster_sharded: a rank produced more cluster records than cap / world");
  return STRGPU_OK;
}

int strgpu_cluster_sharded(strgpu_ctx *ctx, const strgpu_tread *treads, uint32_t n, uint32_t max_n, const strgpu_cluster_params *params,
                           strgpu_bounds *out, uint32_t cap, uint32_t *n_out) {
  if (!ctx || !params || !n_out || (n && !treads) || (cap && !out)) return fail(ctx, STRGPU_ERR_INVALID, "cluster_sharded: null argument");
  if (!ctx->comm) return fail(ctx, STRGPU_ERR_INVALID, "cluster_sharded: strgpu_comm_init has not been called");
  *n_out = 0;
  CU(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure(ctx, ctx->cl_in, (size_t)n * sizeof(strgpu_tread) + 16))) return rc;
  if ((rc = ensure(ctx, ctx->cl_out, (size_t)cap * sizeof(strgpu_bounds) + 16))) return rc;
  cudaStream_t st = ctx->cluster_stream;
  if (n) CU(ctx, cudaMemcpyAsync(ctx->cl_in.p, treads, (size_t)n * sizeof(strgpu_tread), cudaMemcpyHostToDevice, st));
  // the worst-case slot (max_n) always fits; every rank makes the same choice, so the collectives stay matched
  if ((rc = strgpu_cluster_sharded_device(ctx, ctx->cl_in.p, n, max_n, max_n, params, ctx->cl_out.p, cap, ctx->d_cl_n, st))) return rc;
  if ((rc = strgpu_comm_status(ctx, st))) return rc;
  uint32_t produced = 0;
  CU(ctx, cudaMemcpy(&produced, ctx->d_cl_n, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  *n_out = produced;
  const uint32_t take = produced < cap ? produced : cap;
  if (take) CU(ctx, cudaMemcpy(out, ctx->cl_out.p, (size_t)take * sizeof(strgpu_bounds), cudaMemcpyDeviceToHost));
  if (produced > cap) return fail(ctx, STRGPU_ERR_OVERFLOW, "cluster_sharded: %u records produced, capacity %u", produced, cap);
  return STRGPU_OK;
}

}  // extern "C"
