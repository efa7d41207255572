// dtr_nccl.h -- NCCL bound at run time for the sort-first band exchange (dtr_b200_gather_bands,
// dtr_b200_band_barrier).  The library has no link-time NCCL dependency: the first band call
// dlopen()s libnccl.so.2, preferring the copy that is already loaded into the process (a host that
// links NCCL itself, or torch's bundled one), so that one process never runs two NCCL instances.
#pragma once
#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include <cstdlib>
#include <string>

namespace dtr
{

// the handful of NCCL declarations this file needs (nccl.h: ncclUniqueId is 128 opaque bytes passed
// by value, ncclResult_t 0 is success, ncclUint32 = 3, ncclFloat32 = 7, ncclSum = 0)
struct NcclUniqueId
{
	char internal[128];
};
typedef struct ncclComm *NcclComm;
constexpr int NCCL_UINT32 = 3, NCCL_FLOAT32 = 7, NCCL_SUM = 0;

struct NcclApi
{
	bool        ok = false;
	std::string err;
	int (*GetUniqueId)(NcclUniqueId *)                                                           = nullptr;
	int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int)                                      = nullptr;
	int (*CommDestroy)(NcclComm)                                                                 = nullptr;
	int (*GroupStart)()                                                                          = nullptr;
	int (*GroupEnd)()                                                                            = nullptr;
	int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t)                          = nullptr;
	int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t)                                = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t)             = nullptr;
	const char *(*GetErrorString)(int)                                                           = nullptr;
};

inline const NcclApi &nccl_api()
{
	static NcclApi api = [] {
		NcclApi     a;
		void       *h    = nullptr;
		const char *over = getenv("DTR_B200_NCCL_LIB"); // explicit path wins
		if (over && *over) h = dlopen(over, RTLD_NOW | RTLD_GLOBAL);
		if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // already in the process?
		if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!h)
		{
			a.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?");
			return a;
		}
		bool all = true;
		auto sym = [&](const char *name) {
			void *p = dlsym(h, name);
			if (!p)
			{
				all   = false;
				a.err = std::string("NCCL symbol missing: ") + name;
			}
			return p;
		};
		a.GetUniqueId    = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
		a.CommInitRank   = (decltype(a.CommInitRank))sym("ncclCommInitRank");
		a.CommDestroy    = (decltype(a.CommDestroy))sym("ncclCommDestroy");
		a.GroupStart     = (decltype(a.GroupStart))sym("ncclGroupStart");
		a.GroupEnd       = (decltype(a.GroupEnd))sym("ncclGroupEnd");
		a.Send           = (decltype(a.Send))sym("ncclSend");
		a.Recv           = (decltype(a.Recv))sym("ncclRecv");
		a.AllReduce      = (decltype(a.AllReduce))sym("ncclAllReduce");
		a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
		a.ok             = all;
		return a;
	}();
	return api;
}

} // namespace dtr
