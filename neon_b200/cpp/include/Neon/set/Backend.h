// Backend.h — the device set with its stream sets and events; ONE host process drives every GPU, as in the reference.
//
// Mirrors Neon::Backend (libNeonSet/include/Neon/set/Backend.h:26-302): Backend(devIds, Runtime), mainStreamIdx, sync /
// syncAll (:230-261), setAvailableStreamSet (libNeonSet/src/set/Backend.cpp:357-375 — here it only ever GROWS, SURVEY.md
// fact 5), peer access between every device pair (libNeonSet/src/set/DevSet.cpp:80-100).
// Differences by design: launches are asynchronous C-ABI calls, so the host loops over the devices sequentially instead
// of spawning one OpenMP thread per GPU (DevSet.h:372-391); cross-device ordering uses CUDA events, never host syncs.
// An oversubscribed list ({0,0,0}: several partitions on one GPU, as libNeonDomain/tests/domain-halos/src/runHelper.h:64-67
// does) is accepted and is how the multi-partition paths are tested on a single device.
#pragma once

#include <cuda_runtime_api.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "Neon/Neon.h"
#include "Neon/Report.h"

namespace Neon {

#define NEON_CUDA_CHECK(call)                                                                   \
    do {                                                                                        \
        cudaError_t neonCudaErr_ = (call);                                                      \
        if (neonCudaErr_ != cudaSuccess) {                                                      \
            Neon::NeonException neonExc_(#call);                                                \
            neonExc_ << cudaGetErrorName(neonCudaErr_) << ": " << cudaGetErrorString(neonCudaErr_); \
            NEON_THROW(neonExc_);                                                               \
        }                                                                                       \
    } while (0)

struct SetIdx
{
    int idx = 0;
    SetIdx() = default;
    SetIdx(int i) : idx(i) {}
    operator int() const { return idx; }
};

class Backend
{
   public:
    static constexpr int mainStreamIdx = 0;

    Backend() : mS(std::make_shared<State>()) {}
    Backend(const std::vector<int>& devIds, Runtime runtime) : mS(std::make_shared<State>())
    {
        if (devIds.empty()) {
            NeonException e("Backend");
            e << "empty device list";
            NEON_THROW(e);
        }
        mS->devIds = devIds;
        mS->runtime = runtime;
        mS->streams.resize(devIds.size());
        if (runtime == Runtime::stream) {
            const int n = nlbm_device_count();
            if (n <= 0) {
                NeonException e("Backend");
                e << "Runtime::stream needs a CUDA device and there is no CPU fallback (" << nlbm_last_error() << ")";
                NEON_THROW(e);
            }
            for (int id : devIds) {
                if (id < 0 || id >= n) {
                    NeonException e("Backend");
                    e << "invalid CUDA device id " << id << " (" << n << " visible)";
                    NEON_THROW(e);
                }
            }
            for (size_t a = 0; a < devIds.size(); ++a) {
                NEON_CUDA_CHECK(cudaSetDevice(devIds[a]));
                for (size_t b = 0; b < devIds.size(); ++b) {
                    if (devIds[a] != devIds[b]) {
                        detail::check(nlbm_enable_peer_access(devIds[b]), "nlbm_enable_peer_access");
                    }
                }
            }
        }
        setAvailableStreamSet(1);
    }

    /* the slice of Neon::set::DevSet the benchmark-level code asks for (DevSet.h): how many devices, which ids */
    struct DevSetView
    {
        const Backend* bk;
        int            setCardinality() const { return bk->getDeviceCount(); }
        int            devId(int setIdx) const { return bk->devId(setIdx); }
    };
    DevSetView devSet() const { return DevSetView{this}; }

    int     getDeviceCount() const { return int(mS->devIds.size()); }
    int     devId(int setIdx) const { return mS->devIds.at(setIdx); }
    Runtime runtime() const { return mS->runtime; }
    const std::vector<int>& devIds() const { return mS->devIds; }

    void setDevice(int setIdx) const
    {
        if (mS->runtime == Runtime::stream) {
            NEON_CUDA_CHECK(cudaSetDevice(mS->devIds.at(setIdx)));
        }
    }

    void setAvailableStreamSet(int nStreamSets) const
    {
        if (mS->runtime != Runtime::stream) {
            mS->nStreamSets = std::max(mS->nStreamSets, nStreamSets);
            return;
        }
        for (size_t d = 0; d < mS->devIds.size(); ++d) {
            while (int(mS->streams[d].size()) < nStreamSets) {
                setDevice(int(d));
                cudaStream_t s = nullptr;
                NEON_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
                mS->streams[d].push_back(s);
            }
        }
        mS->nStreamSets = std::max(mS->nStreamSets, nStreamSets);
    }
    int getStreamSetCount() const { return mS->nStreamSets; }

    /* the CUDA stream `streamIdx` of device `setIdx` (Backend::streamSet(streamIdx)[setIdx], Backend.h:174) */
    cudaStream_t stream(int setIdx, int streamIdx) const
    {
        if (mS->runtime != Runtime::stream) {
            return nullptr;
        }
        setAvailableStreamSet(streamIdx + 1);
        return mS->streams.at(setIdx).at(streamIdx);
    }

    /* events are created once per use site and owned by the backend */
    cudaEvent_t newEvent(int setIdx) const
    {
        if (mS->runtime != Runtime::stream) {
            return nullptr;
        }
        setDevice(setIdx);
        cudaEvent_t e = nullptr;
        NEON_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        mS->events.push_back({mS->devIds[setIdx], e});
        return e;
    }

    void sync(int streamIdx = mainStreamIdx) const
    {
        if (mS->runtime != Runtime::stream) {
            return;
        }
        for (size_t d = 0; d < mS->devIds.size(); ++d) {
            if (streamIdx < int(mS->streams[d].size())) {
                setDevice(int(d));
                NEON_CUDA_CHECK(cudaStreamSynchronize(mS->streams[d][streamIdx]));
            }
        }
    }
    void sync(SetIdx setIdx, int streamIdx) const
    {
        if (mS->runtime == Runtime::stream) {
            setDevice(setIdx);
            NEON_CUDA_CHECK(cudaStreamSynchronize(stream(setIdx, streamIdx)));
        }
    }
    void syncAll() const
    {
        for (int s = 0; s < mS->nStreamSets; ++s) {
            sync(s);
        }
    }

    std::string toString() const
    {
        std::ostringstream s;
        s << "Backend(" << (mS->runtime == Runtime::stream ? "stream" : mS->runtime == Runtime::openmp ? "openmp" : "none") << ", devices";
        for (int id : mS->devIds) {
            s << " " << id;
        }
        s << ")";
        return s.str();
    }

    void toReport(Report& report) const
    {
        Report sub;
        sub.addMember("Runtime", std::string(mS->runtime == Runtime::stream ? "stream" : "openmp"));
        sub.addMember("nDevices", int(mS->devIds.size()));
        sub.addMember("Devices", mS->devIds);
        if (mS->runtime == Runtime::stream) {
            std::vector<std::string> names;
            for (int id : mS->devIds) {
                cudaDeviceProp p{};
                if (cudaGetDeviceProperties(&p, id) == cudaSuccess) {
                    names.emplace_back(p.name);
                }
            }
            std::string all;
            for (auto& n : names) {
                all += (all.empty() ? "" : "; ") + n;
            }
            sub.addMember("DeviceNames", all);
        }
        sub.addMember("KernelLibrary", std::string("libneon_lbm.so (sm_100a), ABI ") + std::to_string(nlbm_abi_version()));
        report.addSubdoc("Backend", sub);
    }

   private:
    struct State
    {
        std::vector<int>                         devIds;
        Runtime                                  runtime = Runtime::none;
        std::vector<std::vector<cudaStream_t>>   streams;
        std::vector<std::pair<int, cudaEvent_t>> events;
        int                                      nStreamSets = 0;
        ~State()
        {
            for (auto& [dev, e] : events) {
                cudaSetDevice(dev);
                cudaEventDestroy(e);
            }
            for (size_t d = 0; d < streams.size(); ++d) {
                for (cudaStream_t s : streams[d]) {
                    cudaSetDevice(devIds[d]);
                    cudaStreamDestroy(s);
                }
            }
        }
    };
    std::shared_ptr<State> mS;
};

}  // namespace Neon
