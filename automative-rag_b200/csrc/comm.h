// comm.h — host interface of the peer-memory exchange (comm.cu): wire-block export / mapping and the three
// collectives of the sharded stages.  Return codes follow rag_b200.h (0 ok, -1 invalid argument,
// -2 unsupported, -3 CUDA error, -6 peer mapping failed); `err` receives the text.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace rs {

constexpr int kCommMaxWorld = 8;  // the GPUs of one NVSwitch box

struct CommState;
CommState* comm_create(int device, int num_sms);
void comm_destroy(CommState* s);
bool comm_is_open(const CommState* s);
int comm_world(const CommState* s);
int comm_rank(const CommState* s);
size_t comm_slot_bytes(const CommState* s);

int comm_export(CommState* s, int world, int rank, size_t slot_bytes, void* out_blob128, std::string* err);
int comm_open(CommState* s, const void* blobs, std::string* err);
void comm_close(CommState* s);

int comm_allgather_topk(CommState* s, const float* loc_scores, const int64_t* loc_ids, int nq, int k_in, int k_out,
                        float* out_scores, int64_t* out_ids, cudaStream_t stream, std::string* err);
int comm_allgather(CommState* s, const void* local, size_t bytes, void* out, cudaStream_t stream, std::string* err);
int comm_allreduce_max(CommState* s, const float* local, size_t n, float* out, cudaStream_t stream, std::string* err);

}  // namespace rs
