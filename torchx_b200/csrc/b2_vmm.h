// b2_vmm.h — host side of the symmetric arena when it is built from CUDA virtual-memory-management objects:
//   * one physical allocation per rank (cuMemCreate, shareable as a POSIX file descriptor),
//   * every rank maps every peer's allocation (cuMemImportFromShareableHandle + cuMemMap) - the NVSwitch P2P view,
//   * one MULTICAST object spanning all ranks' allocations (cuMulticastCreate / AddDevice / BindMem), mapped once per
//     rank: stores to it are replicated by the switch, multimem.ld_reduce from it is summed by the switch (NVLS).
// File descriptors travel between the one-process-per-GPU workers over abstract-namespace unix datagram sockets with
// SCM_RIGHTS, named after the shm control block of the rendezvous.  The driver API is reached through
// cudaGetDriverEntryPoint, so the library does not link libcuda.
//
// Replaces, on the reference path, what ncclCommInitRank sets up behind torch.distributed.init_process_group("nccl")
// (torchx/distributed/__init__.py:217-222): NCCL's own NVLS transport does the same multicast dance internally.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <poll.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <string>

namespace vmm {

struct Driver {
  bool ok = false;
  std::string why;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
};

template <typename Fn>
static bool load_sym(const char* name, Fn* out, std::string* why) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  const cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    *why = std::string("driver entry point ") + name + " unavailable";
    return false;
  }
  *out = reinterpret_cast<Fn>(p);
  return true;
}

inline const Driver& driver() {
  static const Driver d = [] {
    Driver x;
    bool ok = true;
#define B2_SYM(field, name) ok = ok && load_sym(name, &x.field, &x.why)
    B2_SYM(DeviceGet, "cuDeviceGet");
    B2_SYM(DeviceGetAttribute, "cuDeviceGetAttribute");
    B2_SYM(GetErrorString, "cuGetErrorString");
    B2_SYM(MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
    B2_SYM(MemCreate, "cuMemCreate");
    B2_SYM(MemRelease, "cuMemRelease");
    B2_SYM(MemAddressReserve, "cuMemAddressReserve");
    B2_SYM(MemAddressFree, "cuMemAddressFree");
    B2_SYM(MemMap, "cuMemMap");
    B2_SYM(MemUnmap, "cuMemUnmap");
    B2_SYM(MemSetAccess, "cuMemSetAccess");
    B2_SYM(MemExportToShareableHandle, "cuMemExportToShareableHandle");
    B2_SYM(MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
    B2_SYM(MulticastCreate, "cuMulticastCreate");
    B2_SYM(MulticastAddDevice, "cuMulticastAddDevice");
    B2_SYM(MulticastBindMem, "cuMulticastBindMem");
    B2_SYM(MulticastUnbind, "cuMulticastUnbind");
    B2_SYM(MulticastGetGranularity, "cuMulticastGetGranularity");
#undef B2_SYM
    x.ok = ok;
    return x;
  }();
  return d;
}

inline std::string errstr(CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString && driver().GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "CUresult " + std::to_string(static_cast<int>(r));
}

struct Caps {
  bool vmm = false;        // cuMemCreate + POSIX-fd export usable on this device
  bool multicast = false;  // CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
};

inline Caps caps(int device) {
  Caps c;
  const Driver& d = driver();
  if (!d.ok) return c;
  CUdevice dev;
  if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return c;
  int a = 0, b = 0, m = 0;
  d.DeviceGetAttribute(&a, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
  d.DeviceGetAttribute(&b, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
  d.DeviceGetAttribute(&m, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  c.vmm = a != 0 && b != 0;
  c.multicast = c.vmm && m != 0;
  return c;
}

inline CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

inline CUmulticastObjectProp mc_prop(int world, size_t bytes) {
  CUmulticastObjectProp p;
  memset(&p, 0, sizeof(p));
  p.numDevices = static_cast<unsigned>(world);
  p.size = bytes;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

// Size every rank must use for an arena of at least `want` bytes: a multiple of the allocation granularity and, when
// multicast is in play, of the multicast granularity (identical on identical GPUs, so all ranks compute the same value).
inline size_t arena_granularity(int device, int world, bool multicast) {
  const Driver& d = driver();
  size_t g = 2u << 20;
  const CUmemAllocationProp p = alloc_prop(device);
  size_t a = 0;
  if (d.MemGetAllocationGranularity(&a, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && a > g) g = a;
  if (multicast) {
    const CUmulticastObjectProp mp = mc_prop(world, g);
    size_t m = 0;
    if (d.MulticastGetGranularity(&m, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && m > g) g = m;
  }
  return g;
}

// One mapping of a physical allocation (own or imported) or of the multicast object into this process.
struct Mapping {
  CUmemGenericAllocationHandle handle = 0;
  CUdeviceptr va = 0;
  size_t bytes = 0;
  bool mapped = false;
};

inline std::string map_handle(Mapping* m, size_t bytes, size_t align, const int* devices, int ndev) {
  const Driver& d = driver();
  CUresult r = d.MemAddressReserve(&m->va, bytes, align, 0, 0);
  if (r != CUDA_SUCCESS) return "cuMemAddressReserve: " + errstr(r);
  m->bytes = bytes;
  r = d.MemMap(m->va, bytes, 0, m->handle, 0);
  if (r != CUDA_SUCCESS) {
    d.MemAddressFree(m->va, bytes);
    m->va = 0;
    return "cuMemMap: " + errstr(r);
  }
  m->mapped = true;
  CUmemAccessDesc acc[16];
  for (int i = 0; i < ndev; ++i) {
    memset(&acc[i], 0, sizeof(acc[i]));
    acc[i].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[i].location.id = devices[i];
    acc[i].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  r = d.MemSetAccess(m->va, bytes, acc, static_cast<size_t>(ndev));
  if (r != CUDA_SUCCESS) return "cuMemSetAccess: " + errstr(r);
  return "";
}

inline void unmap_release(Mapping* m) {
  const Driver& d = driver();
  if (m->mapped) d.MemUnmap(m->va, m->bytes);
  if (m->va) d.MemAddressFree(m->va, m->bytes);
  if (m->handle) d.MemRelease(m->handle);
  *m = Mapping();
}

// ---- fd passing ------------------------------------------------------------------------------------------------
struct FdMsg {
  int src_rank;
  int kind;  // 0 = arena of src_rank, 1 = multicast object
};

inline void sock_addr(sockaddr_un* a, socklen_t* len, const std::string& base, int rank) {
  memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  // abstract namespace (leading NUL): no filesystem entry, vanishes with the socket
  const std::string name = base + ".r" + std::to_string(rank);
  const size_t n = name.size() < sizeof(a->sun_path) - 2 ? name.size() : sizeof(a->sun_path) - 2;
  memcpy(a->sun_path + 1, name.data(), n);
  *len = static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + n);
}

inline int sock_open(const std::string& base, int rank, std::string* why) {
  const int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (s < 0) {
    *why = std::string("socket: ") + strerror(errno);
    return -1;
  }
  sockaddr_un a;
  socklen_t len;
  sock_addr(&a, &len, base, rank);
  if (bind(s, reinterpret_cast<sockaddr*>(&a), len) != 0) {
    *why = std::string("bind(abstract unix socket): ") + strerror(errno);
    close(s);
    return -1;
  }
  return s;
}

inline bool send_fd(int sock, const std::string& base, int dst_rank, int fd, const FdMsg& msg, std::string* why) {
  sockaddr_un a;
  socklen_t len;
  sock_addr(&a, &len, base, dst_rank);
  FdMsg payload = msg;
  iovec iov = {&payload, sizeof(payload)};
  alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  memset(ctl, 0, sizeof(ctl));
  msghdr mh;
  memset(&mh, 0, sizeof(mh));
  mh.msg_name = &a;
  mh.msg_namelen = len;
  mh.msg_iov = &iov;
  mh.msg_iovlen = 1;
  mh.msg_control = ctl;
  mh.msg_controllen = sizeof(ctl);
  cmsghdr* cm = CMSG_FIRSTHDR(&mh);
  cm->cmsg_level = SOL_SOCKET;
  cm->cmsg_type = SCM_RIGHTS;
  cm->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  for (int attempt = 0; attempt < 2000; ++attempt) {
    if (sendmsg(sock, &mh, 0) == static_cast<ssize_t>(sizeof(payload))) return true;
    if (errno != ECONNREFUSED && errno != ENOENT && errno != EAGAIN && errno != ENOBUFS) break;
    usleep(1000);  // the peer has not bound its socket yet / its queue is full
  }
  *why = std::string("sendmsg(fd to rank ") + std::to_string(dst_rank) + "): " + strerror(errno);
  return false;
}

// Receives one (fd, msg); returns the fd or -1 (timeout / error).
inline int recv_fd(int sock, FdMsg* msg, int timeout_ms, std::string* why) {
  pollfd p = {sock, POLLIN, 0};
  const int pr = poll(&p, 1, timeout_ms);
  if (pr <= 0) {
    *why = pr == 0 ? "timed out waiting for a peer's file descriptor" : std::string("poll: ") + strerror(errno);
    return -1;
  }
  iovec iov = {msg, sizeof(*msg)};
  alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  msghdr mh;
  memset(&mh, 0, sizeof(mh));
  mh.msg_iov = &iov;
  mh.msg_iovlen = 1;
  mh.msg_control = ctl;
  mh.msg_controllen = sizeof(ctl);
  const ssize_t got = recvmsg(sock, &mh, MSG_CMSG_CLOEXEC);
  if (got != static_cast<ssize_t>(sizeof(*msg))) {
    *why = std::string("recvmsg: ") + (got < 0 ? strerror(errno) : "short message");
    return -1;
  }
  for (cmsghdr* cm = CMSG_FIRSTHDR(&mh); cm; cm = CMSG_NXTHDR(&mh, cm)) {
    if (cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS) {
      int fd = -1;
      memcpy(&fd, CMSG_DATA(cm), sizeof(int));
      return fd;
    }
  }
  *why = "message carried no file descriptor";
  return -1;
}

}  // namespace vmm
