// Peer-memory plumbing for the sharded (one process per GPU) schedule.
//
// The fused re-partition of the four-index transform stores tiles of step 2 straight into the
// destination GPU's buffer (qs_quarter_transform_scatter).  Those buffers are plain cudaMalloc
// allocations exported with CUDA IPC; a rank opens its peers' handles once and keeps the mapped
// pointers.  NVLink/NVSwitch carries the stores; no staging copy and no NCCL call is on the data path.
#include "common.cuh"

extern "C" int qs_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int qs_ipc_alloc(int64_t bytes, void** dev_ptr, void* host_handle) {
    QS_REQUIRE(bytes > 0 && dev_ptr && host_handle, "qs_ipc_alloc: bad arguments");
    void* ptr = nullptr;
    QS_CUDA(cudaMalloc(&ptr, (size_t)bytes));
    cudaIpcMemHandle_t handle;
    cudaError_t e = cudaIpcGetMemHandle(&handle, ptr);
    if (e != cudaSuccess) {
        cudaFree(ptr);
        qs_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return QS_ERR_CUDA;
    }
    memcpy(host_handle, &handle, sizeof(handle));
    *dev_ptr = ptr;
    return QS_OK;
}

extern "C" int qs_ipc_open(const void* host_handle, void** dev_ptr) {
    QS_REQUIRE(host_handle && dev_ptr, "qs_ipc_open: bad arguments");
    cudaIpcMemHandle_t handle;
    memcpy(&handle, host_handle, sizeof(handle));
    void* ptr = nullptr;
    QS_CUDA(cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = ptr;
    return QS_OK;
}

extern "C" int qs_ipc_close(void* dev_ptr) {
    QS_REQUIRE(dev_ptr, "qs_ipc_close: null pointer");
    QS_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return QS_OK;
}

extern "C" int qs_ipc_free(void* dev_ptr) {
    QS_REQUIRE(dev_ptr, "qs_ipc_free: null pointer");
    QS_CUDA(cudaFree(dev_ptr));
    return QS_OK;
}
