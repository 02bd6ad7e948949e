// NCCL plumbing behind fb_comm_* (include/flair_b200.h): the two collectives of the sharded zone loop
//  -- the sum of the per-rank confusion matrices and a byte gather of per-rank map rows to
// the writer rank. libnccl.so.2 is resolved with dlopen at the first call (the copy torch already mapped into the
// process when there is one, the system library otherwise), so the library itself has no link-time dependency on
// NCCL and a single-GPU host never touches it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace fb {

struct Comm;  // one NCCL communicator: rank `rank` of `world`, bound to the device current at comm_init

constexpr int kCommIdBytes = 128;  // sizeof(ncclUniqueId)

// 0 on success; a negative library code with *err set otherwise.
int comm_unique_id(uint8_t* id128, std::string* err);
int comm_init(Comm** out, const uint8_t* id128, int rank, int world, std::string* err);
void comm_destroy(Comm* c);
int comm_rank(const Comm* c);
int comm_world(const Comm* c);
// in-place sum over ranks of n int64 values (device buffer), enqueued on `stream`
int comm_allreduce_i64(Comm* c, long long* buf_dev, size_t n, cudaStream_t stream, std::string* err);
// rank r's send_bytes bytes land at recv_dev + offsets[r] on `root` (offsets / counts: host arrays of `world` entries,
// identical on every rank; recv_dev is only read on root). Device buffers, enqueued on `stream`.
int comm_gather_bytes(Comm* c, const void* send_dev, long long send_bytes, void* recv_dev, const long long* offsets,
                      const long long* counts, int root, cudaStream_t stream, std::string* err);

}  // namespace fb
