// NCCL bound at run time, on first use: single-GPU contexts (every LIFE example, the 1-GPU benchmark) never touch it, and mapping
// libnccl.so (hundreds of MB of device code) costs up to seconds of start-up on a cold box.  The macros below redirect the
// handful of NCCL entry points liblife_b200 uses to function pointers resolved with dlopen("libnccl.so.2") — which also returns
// the copy a host process (e.g. PyTorch) has already loaded, so there is never a second NCCL in the process.
#pragma once
#include <nccl.h>

namespace life {

struct NcclApi {
	decltype(&ncclGetUniqueId) GetUniqueId;
	decltype(&ncclCommInitRank) CommInitRank;
	decltype(&ncclCommDestroy) CommDestroy;
	decltype(&ncclCommGetAsyncError) CommGetAsyncError;
	decltype(&ncclGetErrorString) GetErrorString;
	decltype(&ncclGroupStart) GroupStart;
	decltype(&ncclGroupEnd) GroupEnd;
	decltype(&ncclSend) Send;
	decltype(&ncclRecv) Recv;
	decltype(&ncclAllReduce) AllReduce;
};

// resolved on first call; if libnccl.so.2 cannot be loaded every entry returns ncclSystemError and GetErrorString says why
const NcclApi &nccl_api();

}  // namespace life

#ifndef LIFE_NCCL_NO_REDIRECT
#define ncclGetUniqueId life::nccl_api().GetUniqueId
#define ncclCommInitRank life::nccl_api().CommInitRank
#define ncclCommDestroy life::nccl_api().CommDestroy
#define ncclCommGetAsyncError life::nccl_api().CommGetAsyncError
#define ncclGetErrorString life::nccl_api().GetErrorString
#define ncclGroupStart life::nccl_api().GroupStart
#define ncclGroupEnd life::nccl_api().GroupEnd
#define ncclSend life::nccl_api().Send
#define ncclRecv life::nccl_api().Recv
#define ncclAllReduce life::nccl_api().AllReduce
#endif
