// Collectives of the multi-GPU partitioner (one process per GPU).  Two back-ends behind one
// interface: NCCL over NVLink 5 / NVSwitch, resolved at run time from the libnccl already loaded
// in the process (torch's bundled copy) so the single-GPU path has no NCCL dependency; and a
// host-callback back-end (the launcher supplies all-reduce / send-recv, e.g. gloo) that lets
// several ranks share one GPU in tests.  The reference has no counterpart (it is single GPU).
#ifndef PGN_COMM_H
#define PGN_COMM_H

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pagani_b200.h"

namespace pgn {

enum ReduceOp { kOpSum = 0, kOpMax = 1 };

// One contiguous run of regions that moves from `src` to `dst` during a rebalance.
struct Transfer {
  int src, dst;
  size_t src_offset;  // first region, in src's local list
  size_t dst_offset;  // position in dst's new local list
  size_t count;
};

// Order-preserving repartition of the concatenated (rank-major) region list into near-equal
// contiguous slices.  Pure host logic (unit-tested on CPU): every rank computes the same plan.
void plan_rebalance(const std::vector<size_t>& counts, std::vector<size_t>* new_counts,
                    std::vector<Transfer>* transfers);

class Comm {
public:
  virtual ~Comm() {}
  int nranks = 1;
  int rank = 0;
  // in place on a device buffer, ordered on `stream`
  virtual int allreduce(double* dev, int n, ReduceOp op, cudaStream_t stream) = 0;
  // exchange packed device buffers according to `plan` (only entries touching this rank are used);
  // send_base/recv_base are device pointers, offsets in doubles are given per transfer.
  virtual int exchange(const std::vector<Transfer>& plan, const double* send_base,
                       const std::vector<size_t>& send_off, double* recv_base,
                       const std::vector<size_t>& recv_off, size_t doubles_per_region,
                       cudaStream_t stream) = 0;
  virtual const char* name() const = 0;
};

Comm* make_nccl_comm(const uint8_t id[PGN_NCCL_ID_BYTES], int nranks, int rank, std::string* err);
int nccl_unique_id(uint8_t id[PGN_NCCL_ID_BYTES], std::string* err);

Comm* make_host_comm(int nranks, int rank, pgn_allreduce_fn ar, pgn_alltoallv_fn a2a, void* user);

}  // namespace pgn
#endif
