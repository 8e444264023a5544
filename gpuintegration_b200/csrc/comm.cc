#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>

namespace pgn {

void
plan_rebalance(const std::vector<size_t>& counts, std::vector<size_t>* new_counts,
               std::vector<Transfer>* transfers)
{
  const int w = static_cast<int>(counts.size());
  size_t total = 0;
  for (size_t c : counts)
    total += c;
  new_counts->assign(w, 0);
  transfers->clear();
  if (w == 0)
    return;
  // rank s owns global positions [q[s], q[s+1]): the first (total % w) ranks get one extra region
  std::vector<size_t> q(w + 1, 0), p(w + 1, 0);
  const size_t each = total / w, extra = total % w;
  for (int s = 0; s < w; ++s) {
    (*new_counts)[s] = each + (static_cast<size_t>(s) < extra ? 1 : 0);
    q[s + 1] = q[s] + (*new_counts)[s];
    p[s + 1] = p[s] + counts[s];
  }
  for (int src = 0; src < w; ++src) {
    for (int dst = 0; dst < w; ++dst) {
      const size_t lo = p[src] > q[dst] ? p[src] : q[dst];
      const size_t hi = p[src + 1] < q[dst + 1] ? p[src + 1] : q[dst + 1];
      if (hi > lo) {
        Transfer t;
        t.src = src;
        t.dst = dst;
        t.src_offset = lo - p[src];
        t.dst_offset = lo - q[dst];
        t.count = hi - lo;
        transfers->push_back(t);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ------------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi*
nccl_api(std::string* err)
{
  static NcclApi api;
  static bool tried = false;
  if (api.handle)
    return &api;
  if (tried) {
    if (err)
      *err = "libnccl.so.2 could not be loaded";
    return nullptr;
  }
  tried = true;
  // RTLD_NOLOAD first: reuse the copy torch already mapped (same NCCL as torch.distributed)
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h)
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h)
    h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    if (err)
      *err = std::string("dlopen(libnccl.so.2): ") + dlerror();
    return nullptr;
  }
#define PGN_SYM(field, sym)                                                                        \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym));                                \
  if (!api.field) {                                                                                \
    if (err)                                                                                       \
      *err = std::string("missing NCCL symbol ") + sym;                                            \
    return nullptr;                                                                                \
  }
  PGN_SYM(GetUniqueId, "ncclGetUniqueId")
  PGN_SYM(CommInitRank, "ncclCommInitRank")
  PGN_SYM(CommDestroy, "ncclCommDestroy")
  PGN_SYM(AllReduce, "ncclAllReduce")
  PGN_SYM(Send, "ncclSend")
  PGN_SYM(Recv, "ncclRecv")
  PGN_SYM(GroupStart, "ncclGroupStart")
  PGN_SYM(GroupEnd, "ncclGroupEnd")
  PGN_SYM(GetErrorString, "ncclGetErrorString")
#undef PGN_SYM
  api.handle = h;
  return &api;
}

class NcclComm : public Comm {
public:
  NcclApi* api = nullptr;
  ncclComm_t comm = nullptr;
  ~NcclComm() override
  {
    if (comm)
      api->CommDestroy(comm);
  }
  int
  allreduce(double* dev, int n, ReduceOp op, cudaStream_t stream) override
  {
    const ncclResult_t rc =
      api->AllReduce(dev, dev, static_cast<size_t>(n), ncclDouble, op == kOpSum ? ncclSum : ncclMax,
                     comm, stream);
    return rc == ncclSuccess ? PGN_OK : PGN_ERR_NCCL;
  }
  int
  exchange(const std::vector<Transfer>& plan, const double* send_base,
           const std::vector<size_t>& send_off, double* recv_base,
           const std::vector<size_t>& recv_off, size_t dpr, cudaStream_t stream) override
  {
    if (api->GroupStart() != ncclSuccess)
      return PGN_ERR_NCCL;
    bool ok = true;  // a failed call must not leave the group open: GroupEnd runs in every case
    for (size_t k = 0; k < plan.size() && ok; ++k) {
      const Transfer& t = plan[k];
      if (t.src == t.dst)
        continue;
      if (t.src == rank)
        ok = api->Send(send_base + send_off[k], t.count * dpr, ncclDouble, t.dst, comm, stream) == ncclSuccess;
      if (ok && t.dst == rank)
        ok = api->Recv(recv_base + recv_off[k], t.count * dpr, ncclDouble, t.src, comm, stream) == ncclSuccess;
    }
    return api->GroupEnd() == ncclSuccess && ok ? PGN_OK : PGN_ERR_NCCL;
  }
  const char* name() const override { return "nccl"; }
};

class HostComm : public Comm {
public:
  pgn_allreduce_fn ar = nullptr;
  pgn_alltoallv_fn a2a = nullptr;
  void* user = nullptr;
  std::vector<double> tmp;
  int
  allreduce(double* dev, int n, ReduceOp op, cudaStream_t stream) override
  {
    tmp.resize(n);
    if (cudaMemcpyAsync(tmp.data(), dev, sizeof(double) * n, cudaMemcpyDeviceToHost, stream) !=
          cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
      return PGN_ERR_CUDA;
    if (ar(user, tmp.data(), n, static_cast<int>(op)) != 0)
      return PGN_ERR_NCCL;
    if (cudaMemcpyAsync(dev, tmp.data(), sizeof(double) * n, cudaMemcpyHostToDevice, stream) !=
        cudaSuccess)
      return PGN_ERR_CUDA;
    return cudaStreamSynchronize(stream) == cudaSuccess ? PGN_OK : PGN_ERR_CUDA;
  }
  int
  exchange(const std::vector<Transfer>& plan, const double* send_base,
           const std::vector<size_t>& send_off, double* recv_base,
           const std::vector<size_t>& recv_off, size_t dpr, cudaStream_t stream) override
  {
    // stage through host memory: one contiguous block per peer, peers in rank order
    std::vector<uint64_t> scount(nranks, 0), rcount(nranks, 0), soff(nranks, 0), roff(nranks, 0);
    for (const Transfer& t : plan) {
      if (t.src == t.dst)
        continue;
      if (t.src == rank)
        scount[t.dst] += t.count * dpr;
      if (t.dst == rank)
        rcount[t.src] += t.count * dpr;
    }
    uint64_t stot = 0, rtot = 0;
    for (int r = 0; r < nranks; ++r) {
      soff[r] = stot;
      stot += scount[r];
      roff[r] = rtot;
      rtot += rcount[r];
    }
    std::vector<double> sbuf(stot), rbuf(rtot);
    std::vector<uint64_t> sfill(soff), rfill(roff);
    for (size_t k = 0; k < plan.size(); ++k) {
      const Transfer& t = plan[k];
      if (t.src == t.dst || t.src != rank)
        continue;
      if (cudaMemcpyAsync(sbuf.data() + sfill[t.dst], send_base + send_off[k],
                          sizeof(double) * t.count * dpr, cudaMemcpyDeviceToHost,
                          stream) != cudaSuccess)
        return PGN_ERR_CUDA;
      sfill[t.dst] += t.count * dpr;
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess)
      return PGN_ERR_CUDA;
    if (a2a(user, sbuf.data(), scount.data(), soff.data(), rbuf.data(), rcount.data(),
            roff.data()) != 0)
      return PGN_ERR_NCCL;
    for (size_t k = 0; k < plan.size(); ++k) {
      const Transfer& t = plan[k];
      if (t.src == t.dst || t.dst != rank)
        continue;
      if (cudaMemcpyAsync(recv_base + recv_off[k], rbuf.data() + rfill[t.src],
                          sizeof(double) * t.count * dpr, cudaMemcpyHostToDevice,
                          stream) != cudaSuccess)
        return PGN_ERR_CUDA;
      rfill[t.src] += t.count * dpr;
    }
    return cudaStreamSynchronize(stream) == cudaSuccess ? PGN_OK : PGN_ERR_CUDA;
  }
  const char* name() const override { return "host"; }
};

}  // namespace

int
nccl_unique_id(uint8_t id[PGN_NCCL_ID_BYTES], std::string* err)
{
  NcclApi* api = nccl_api(err);
  if (!api)
    return PGN_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) <= PGN_NCCL_ID_BYTES, "id buffer too small");
  ncclUniqueId uid;
  const ncclResult_t rc = api->GetUniqueId(&uid);
  if (rc != ncclSuccess) {
    if (err)
      *err = api->GetErrorString(rc);
    return PGN_ERR_NCCL;
  }
  std::memset(id, 0, PGN_NCCL_ID_BYTES);
  std::memcpy(id, &uid, sizeof(uid));
  return PGN_OK;
}

Comm*
make_nccl_comm(const uint8_t id[PGN_NCCL_ID_BYTES], int nranks, int rank, std::string* err)
{
  NcclApi* api = nccl_api(err);
  if (!api)
    return nullptr;
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  NcclComm* c = new NcclComm;
  c->api = api;
  c->nranks = nranks;
  c->rank = rank;
  const ncclResult_t rc = api->CommInitRank(&c->comm, nranks, uid, rank);
  if (rc != ncclSuccess) {
    if (err)
      *err = api->GetErrorString(rc);
    c->comm = nullptr;
    delete c;
    return nullptr;
  }
  return c;
}

Comm*
make_host_comm(int nranks, int rank, pgn_allreduce_fn ar, pgn_alltoallv_fn a2a, void* user)
{
  HostComm* c = new HostComm;
  c->nranks = nranks;
  c->rank = rank;
  c->ar = ar;
  c->a2a = a2a;
  c->user = user;
  return c;
}

}  // namespace pgn
