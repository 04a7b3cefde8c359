/* C ABI of the bagua_b200 native core (reference counterpart: rust/bagua-core/bagua-core-c/src/lib.rs:22-347).
 * Lets non-python runtimes drive the scheduler: create tensors (from raw device pointers), buckets and a backend,
 * attach communication ops, mark tensors ready and wait. Every function returns 0 on success, a negative code on error;
 * bagua_last_error() describes the last failure of the calling thread. Handles are opaque pointers. */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAGUA_API __attribute__((visibility("default")))

typedef struct BaguaTensorC BaguaTensorC;
typedef struct BaguaBucketC BaguaBucketC;
typedef struct BaguaBackendC BaguaBackendC;
typedef struct BaguaPeerCommC BaguaPeerCommC;
typedef void (*bagua_callback_t)(const char* bucket_name, void* user_data);

BAGUA_API const char* bagua_last_error(void);
BAGUA_API const char* bagua_version(void);

/* dtype codes: 0 f32, 1 f16, 2 u8, 3 i64, 4 bf16.  device < 0 = host memory / CPU backend. */
BAGUA_API BaguaTensorC* bagua_tensor_c_create(const char* name, uint64_t data_ptr, int64_t num_elements, int dtype, int device_id);
BAGUA_API void bagua_tensor_c_destroy(BaguaTensorC** tensor);

BAGUA_API BaguaBucketC* bagua_bucket_c_create(BaguaTensorC* const* tensors, size_t n_tensors, const char* name);
BAGUA_API void bagua_bucket_c_destroy(BaguaBucketC** bucket);
BAGUA_API int bagua_bucket_c_append_callback_op(BaguaBucketC* bucket, bagua_callback_t fn, void* user_data);
BAGUA_API int bagua_bucket_c_clear_ops(BaguaBucketC* bucket);

/* NVSwitch communicator: flag_ptrs[p] = address of rank p's (zeroed, symmetric) signal pad mapped in this process. */
BAGUA_API BaguaPeerCommC* bagua_peer_comm_c_create(int rank, int nranks, int device_id, const uint64_t* flag_ptrs, double timeout_s);
BAGUA_API void bagua_peer_comm_c_destroy(BaguaPeerCommC** comm);
BAGUA_API int bagua_peer_comm_c_abort(BaguaPeerCommC* comm);
BAGUA_API size_t bagua_peer_comm_c_signal_pad_bytes(void);

/* In-place fused allreduce over a symmetric buffer: peer_ptrs[p] = rank p's mapping of the buffer, mc_ptr = NVLS multicast
 * alias or 0. variant: 1 two-shot, 2 multimem. average != 0 divides by nranks in the kernel epilogue. */
BAGUA_API int bagua_bucket_c_append_centralized_synchronous_op(BaguaBucketC* bucket, BaguaPeerCommC* comm, const uint64_t* peer_ptrs,
                                                              uint64_t mc_ptr, size_t offset, size_t bytes, int dtype, int average, int variant,
                                                              int nblocks);

BAGUA_API BaguaBackendC* bagua_comm_backend_c_create(size_t schedule_channel_cap, int device_id, uint64_t comm_stream, double watchdog_timeout_s);
BAGUA_API void bagua_comm_backend_c_destroy(BaguaBackendC** backend);
BAGUA_API int bagua_comm_backend_c_register_ordered_buckets(BaguaBackendC* backend, BaguaBucketC* const* buckets, size_t n_buckets);
BAGUA_API int bagua_comm_backend_c_mark_communication_ready(BaguaBackendC* backend, BaguaTensorC* tensor, uint64_t ready_cuda_event);
BAGUA_API int bagua_comm_backend_c_wait_pending_comm_ops(BaguaBackendC* backend, uint64_t consumer_stream, int host_sync);

#ifdef __cplusplus
}
#endif
