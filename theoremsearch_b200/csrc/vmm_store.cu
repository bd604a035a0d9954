// Growable device buffer for the corpus rows: a large virtual address range is reserved once and physical memory is
// mapped into it chunk by chunk (cuMemAddressReserve / cuMemCreate / cuMemMap, reached through the runtime's driver
// entry points — no libcuda link dependency). Growing the store therefore never copies a row and never needs the old and
// the new allocation resident together: a 100 GB shard on a 180 GB B200 can still grow. The pointer is ordinary global
// memory to every kernel (1-D bulk TMA, tensor maps, loads). If the driver entry points are unavailable the store
// falls back to cudaMalloc + copy-on-grow.
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "ts_common.cuh"

namespace ts {

struct VmmApi {
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*afree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*gran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

static VmmApi vmm_resolve() {
    VmmApi api;
    auto get = [](const char* name, void** fn) {
        cudaDriverEntryPointQueryResult q;
        return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess &&
               *fn != nullptr;
    };
    api.ok = get("cuMemAddressReserve", (void**)&api.reserve) && get("cuMemAddressFree", (void**)&api.afree) &&
             get("cuMemCreate", (void**)&api.create) && get("cuMemRelease", (void**)&api.release) &&
             get("cuMemMap", (void**)&api.map) && get("cuMemUnmap", (void**)&api.unmap) &&
             get("cuMemSetAccess", (void**)&api.access) && get("cuMemGetAllocationGranularity", (void**)&api.gran);
    if (!api.ok) cudaGetLastError();
    return api;
}

static const VmmApi& vmm_api() {
    static const VmmApi api = vmm_resolve();   // resolved once; initialisation of a local static is thread-safe
    return api;
}

struct RowStore {
    int device = 0;
    bool vmm = false;
    void* base = nullptr;
    size_t va_bytes = 0;      // reserved address range (vmm)
    size_t mapped = 0;        // bytes backed by physical memory (vmm) / allocated (fallback)
    size_t gran = 0;
    std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;   // (handle, bytes), in address order
};

static CUmemAllocationProp vmm_prop(int device) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    return prop;
}

// map `bytes` (rounded up to the granularity) more physical memory at the end of the mapped range
static int vmm_extend(RowStore* st, size_t bytes) {
    const VmmApi& api = vmm_api();
    const size_t add = (bytes + st->gran - 1) / st->gran * st->gran;
    TS_REQUIRE(st->mapped + add <= st->va_bytes, TS_ERR_CAPACITY, "row store: %zu bytes exceed the reserved address range (%zu)",
               st->mapped + add, st->va_bytes);
    const CUmemAllocationProp prop = vmm_prop(st->device);
    CUmemGenericAllocationHandle h;
    CUresult r = api.create(&h, add, &prop, 0);
    if (r != CUDA_SUCCESS) {
        set_error("row store: cuMemCreate(%zu bytes) failed (CUresult %d)", add, (int)r);
        return r == CUDA_ERROR_OUT_OF_MEMORY ? TS_ERR_OOM : TS_ERR_CUDA;
    }
    const CUdeviceptr at = (CUdeviceptr)st->base + st->mapped;
    r = api.map(at, add, 0, h, 0);
    if (r == CUDA_SUCCESS) {
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = st->device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = api.access(at, add, &acc, 1);
        if (r != CUDA_SUCCESS) api.unmap(at, add);
    }
    if (r != CUDA_SUCCESS) {
        api.release(h);
        set_error("row store: cuMemMap / cuMemSetAccess failed (CUresult %d)", (int)r);
        return TS_ERR_CUDA;
    }
    st->chunks.emplace_back(h, add);
    st->mapped += add;
    return TS_OK;
}

int row_store_create(RowStore** out, int device, size_t bytes, size_t row_bytes) {
    RowStore* st = new RowStore();
    st->device = device;
    cudaFree(nullptr);        // the driver entry points below need the device's primary context to be current
    const VmmApi& api = vmm_api();
    if (api.ok && !tunables().store_no_vmm) {
        const CUmemAllocationProp prop = vmm_prop(device);
        size_t gran = 0, free_b = 0, total_b = 0;
        if (api.gran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && gran > 0 &&
            cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            // address range: every row the 32-bit row keys can name, but no more than the device could ever hold
            size_t va = std::min<size_t>((size_t)0xFFFFFFFEull * row_bytes, total_b);
            va = std::max(va, bytes);
            va = (va + gran - 1) / gran * gran;
            CUdeviceptr p = 0;
            if (api.reserve(&p, va, 0, 0, 0) == CUDA_SUCCESS) {
                st->vmm = true;
                st->base = (void*)p;
                st->va_bytes = va;
                st->gran = gran;
                int rc = vmm_extend(st, std::max<size_t>(bytes, 1));
                if (rc != TS_OK) {
                    api.afree(p, va);
                    delete st;
                    return rc;
                }
                *out = st;
                return TS_OK;
            }
        }
        cudaGetLastError();
    }
    cudaError_t e = cudaMalloc(&st->base, std::max<size_t>(bytes, 1));
    if (e != cudaSuccess) {
        set_error("row store: cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        delete st;
        cudaGetLastError();
        return TS_ERR_OOM;
    }
    st->mapped = std::max<size_t>(bytes, 1);
    *out = st;
    return TS_OK;
}

void row_store_destroy(RowStore* st) {
    if (!st) return;
    if (st->vmm) {
        cudaDeviceSynchronize();   // cudaFree waits for the device by itself, cuMemUnmap does not
        const VmmApi& api = vmm_api();
        CUdeviceptr at = (CUdeviceptr)st->base;
        for (auto& c : st->chunks) {
            api.unmap(at, c.second);
            api.release(c.first);
            at += c.second;
        }
        api.afree((CUdeviceptr)st->base, st->va_bytes);
    } else {
        cudaFree(st->base);
    }
    delete st;
}

void* row_store_ptr(const RowStore* st) { return st ? st->base : nullptr; }
bool row_store_is_vmm(const RowStore* st) { return st && st->vmm; }

// make at least `bytes` usable; `used` bytes hold data (copied across by the fallback path). The device must be idle.
int row_store_reserve(RowStore* st, size_t bytes, size_t used) {
    if (bytes <= st->mapped) return TS_OK;
    if (st->vmm) return vmm_extend(st, bytes - st->mapped);
    void* fresh = nullptr;
    cudaError_t e = cudaMalloc(&fresh, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("row store: cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return TS_ERR_OOM;
    }
    if (used > 0) {
        e = cudaMemcpy(fresh, st->base, used, cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) {
            cudaFree(fresh);
            set_error("row store: device copy failed: %s", cudaGetErrorString(e));
            return TS_ERR_CUDA;
        }
    }
    cudaFree(st->base);
    st->base = fresh;
    st->mapped = bytes;
    return TS_OK;
}

}  // namespace ts
