// vmm.hpp -- index storage that grows in place: one virtual address range per array, physical memory mapped on demand.
//
// The reference's storage is a ConcurrentLinkedQueue that grows one node per append (BruteForceIndex.scala:34-36,48-52); a
// flat device matrix that grows by cudaMalloc + copy + cudaFree needs 2x its size in HBM while it grows (fatal for a
// 100M-row shard on a 180 GB device) and moves every pointer under the feet of concurrent queries.  Here every array
// (rows, ids, norms, shadow) reserves its maximum virtual range once (cuMemAddressReserve) and maps physical chunks behind
// the rows as they arrive (cuMemCreate / cuMemMap / cuMemSetAccess): growing never copies, never doubles the footprint and
// never changes a base pointer, so appends can proceed while queries read the rows published so far.
// The driver entry points come through cudaGetDriverEntryPoint (no link-time dependency on libcuda, like the tensor maps).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <vector>

namespace b200ann {

struct VmApi {
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

inline const VmApi& vm_api() {
    static VmApi api = [] {
        VmApi a;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        a.ok = get("cuMemAddressReserve", (void**)&a.AddressReserve) && get("cuMemAddressFree", (void**)&a.AddressFree) &&
               get("cuMemCreate", (void**)&a.Create) && get("cuMemRelease", (void**)&a.Release) && get("cuMemMap", (void**)&a.Map) &&
               get("cuMemUnmap", (void**)&a.Unmap) && get("cuMemSetAccess", (void**)&a.SetAccess) &&
               get("cuMemGetAllocationGranularity", (void**)&a.GetGranularity);
        (void)cudaGetLastError();
        return a;
    }();
    return api;
}

// One growable array.  `ensure(bytes)` maps more physical memory behind what is mapped already; the new range is reported
// through (*fresh_off, *fresh_len) so that the caller can zero it on its stream.
struct VmArray {
    CUdeviceptr base = 0;
    size_t reserved = 0, mapped = 0, gran = 0;
    int device = 0;
    struct Chunk {
        CUmemGenericAllocationHandle h;
        size_t off, len;
    };
    std::vector<Chunk> chunks;

    void* ptr() const { return reinterpret_cast<void*>(base); }

    // cudaSuccess, cudaErrorMemoryAllocation (no address space / no memory) or cudaErrorNotSupported
    cudaError_t reserve(size_t bytes, int dev) {
        const VmApi& a = vm_api();
        if (!a.ok) return cudaErrorNotSupported;
        device = dev;
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = dev;
        if (a.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return cudaErrorNotSupported;
        size_t want = (bytes + gran - 1) / gran * gran;
        if (a.AddressReserve(&base, want, 0, 0, 0) != CUDA_SUCCESS) {
            base = 0;
            return cudaErrorMemoryAllocation;
        }
        reserved = want;
        return cudaSuccess;
    }

    cudaError_t ensure(size_t bytes, size_t* fresh_off, size_t* fresh_len) {
        *fresh_off = mapped;
        *fresh_len = 0;
        if (bytes <= mapped) return cudaSuccess;
        if (bytes > reserved) return cudaErrorMemoryAllocation;
        const VmApi& a = vm_api();
        const size_t need = (bytes + gran - 1) / gran * gran - mapped;
        // growth policy: at least what is asked, and a quarter of what is mapped (32 MB .. 1 GB) so that a stream of small
        // appends maps O(log n) chunks; fall back to the bare minimum when the device is nearly full
        size_t step = std::max(need, std::min<size_t>(std::max<size_t>(mapped / 4, (size_t)32 << 20), (size_t)1 << 30));
        step = std::min((step + gran - 1) / gran * gran, reserved - mapped);
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device;
        CUmemGenericAllocationHandle h;
        CUresult r = a.Create(&h, step, &prop, 0);
        if (r != CUDA_SUCCESS && step > need) {
            step = need;
            r = a.Create(&h, step, &prop, 0);
        }
        if (r != CUDA_SUCCESS) return cudaErrorMemoryAllocation;
        if (a.Map(base + mapped, step, 0, h, 0) != CUDA_SUCCESS) {
            a.Release(h);
            return cudaErrorMemoryAllocation;
        }
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (a.SetAccess(base + mapped, step, &acc, 1) != CUDA_SUCCESS) {
            a.Unmap(base + mapped, step);
            a.Release(h);
            return cudaErrorMemoryAllocation;
        }
        chunks.push_back({h, mapped, step});
        *fresh_len = step;
        mapped += step;
        return cudaSuccess;
    }

    void release() {
        const VmApi& a = vm_api();
        if (!a.ok) return;
        for (const Chunk& c : chunks) {
            a.Unmap(base + c.off, c.len);
            a.Release(c.h);
        }
        chunks.clear();
        if (base) a.AddressFree(base, reserved);
        base = 0;
        reserved = mapped = 0;
    }
};

}  // namespace b200ann
