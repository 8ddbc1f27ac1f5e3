// Run-time binding of NCCL (see nccl_dyn.h).
#define LIFE_NCCL_NO_REDIRECT
#include "nccl_dyn.h"
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <type_traits>

namespace life {

static std::string g_why;

// stand-ins used when the library is missing: fail loudly, never pretend
static ncclResult_t no_uid(ncclUniqueId *) { return ncclSystemError; }
static ncclResult_t no_init(ncclComm_t *, int, ncclUniqueId, int) { return ncclSystemError; }
static ncclResult_t no_destroy(ncclComm_t) { return ncclSystemError; }
static ncclResult_t no_async(ncclComm_t, ncclResult_t *) { return ncclSystemError; }
static const char *no_str(ncclResult_t) { return g_why.c_str(); }
static ncclResult_t no_group() { return ncclSystemError; }
static ncclResult_t no_send(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclSystemError; }
static ncclResult_t no_recv(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclSystemError; }
static ncclResult_t no_allreduce(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) { return ncclSystemError; }

const NcclApi &nccl_api() {
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		api = NcclApi{no_uid, no_init, no_destroy, no_async, no_str, no_group, no_group, no_send, no_recv, no_allreduce};
		void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!h) {
			const char *e = dlerror();
			g_why = std::string("libnccl.so.2 could not be loaded (") + (e ? e : "unknown reason") + ")";
			return;
		}
		NcclApi a{};
		bool ok = true;
		auto bind = [&](auto &fp, const char *name) {
			void *p = dlsym(h, name);
			if (!p) { ok = false; g_why = std::string("libnccl.so.2 lacks ") + name; }
			fp = reinterpret_cast<std::remove_reference_t<decltype(fp)>>(p);
		};
		bind(a.GetUniqueId, "ncclGetUniqueId");
		bind(a.CommInitRank, "ncclCommInitRank");
		bind(a.CommDestroy, "ncclCommDestroy");
		bind(a.CommGetAsyncError, "ncclCommGetAsyncError");
		bind(a.GetErrorString, "ncclGetErrorString");
		bind(a.GroupStart, "ncclGroupStart");
		bind(a.GroupEnd, "ncclGroupEnd");
		bind(a.Send, "ncclSend");
		bind(a.Recv, "ncclRecv");
		bind(a.AllReduce, "ncclAllReduce");
		if (ok) api = a;
	});
	return api;
}

}  // namespace life
