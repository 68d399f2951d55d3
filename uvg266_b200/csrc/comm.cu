// comm.cu - the one collective of the path: the reference-picture exchange for inter pictures when a picture is
// sharded across GPUs by CTU-row bands. Every GPU reconstructs only its own band; before the next
// picture's motion search every GPU must hold the whole reconstructed picture. That is one broadcast per band, all
// issued as ONE NCCL group (a single fused launch on the stream), in place on the picture buffer - no staging copies,
// bands may differ by a CTU row. All-intra pictures need no communication at all.
//
// One process per GPU. NCCL is bound at run time (dlopen of libnccl.so.2, the one the host process already uses if
// any), so the library loads and every single-GPU entry point works on machines without NCCL; the communicator
// bootstrap id travels through whatever the host application has (bench.py: torch.distributed broadcast; an encoder:
// its own control channel). No NCCL type appears in the C ABI.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>
#include <new>
#include <string.h>

namespace uvgb {
namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId is 128 bytes, results are ints, ncclUint8 = 1)
struct NcclId { char internal[128]; };
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
  int (*GetUniqueId)(NcclId *);
  int (*CommInitRank)(ncclComm_t *, int, NcclId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
  int (*GetVersion)(int *);
  bool ok = false;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl()
{
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return;
  auto sym = [&](const char *n) { return dlsym(h, n); };
  g_nccl.GetUniqueId = reinterpret_cast<int (*)(NcclId *)>(sym("ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<int (*)(ncclComm_t *, int, NcclId, int)>(sym("ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<int (*)(ncclComm_t)>(sym("ncclCommDestroy"));
  g_nccl.Broadcast = reinterpret_cast<int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t)>(sym("ncclBroadcast"));
  g_nccl.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
  g_nccl.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
  g_nccl.GetErrorString = reinterpret_cast<const char *(*)(int)>(sym("ncclGetErrorString"));
  g_nccl.GetVersion = reinterpret_cast<int (*)(int *)>(sym("ncclGetVersion"));
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.Broadcast && g_nccl.GroupStart &&
              g_nccl.GroupEnd && g_nccl.GetErrorString;
}

int need_nccl()
{
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) { set_error("NCCL (libnccl.so.2) is not available in this process: %s", dlerror() ? dlerror() : "symbols missing"); return 1; }
  return 0;
}

#define UVGB_NCCL(expr)                                                                          \
  do {                                                                                           \
    const int r__ = (expr);                                                                      \
    if (r__ != 0) { set_error("%s failed: %s", #expr, g_nccl.GetErrorString(r__)); return 1; }   \
  } while (0)

}  // namespace
}  // namespace uvgb

using namespace uvgb;

struct uvg_b200_comm { ncclComm_t comm; int rank, world; };

// CTU rows [lo, hi) of band `rank` of `world`: balanced, identical on every rank (the rule of bench.py's bands).
extern "C" void uvg_b200_band_rows(int rank, int world, int ctu_rows, int *lo, int *hi)
{
  *lo = (int)((long long)ctu_rows * rank / world);
  *hi = (int)((long long)ctu_rows * (rank + 1) / world);
}

extern "C" int uvg_b200_comm_unique_id(void *id_out)
{
  if (ensure_init() || need_nccl()) return 1;
  if (!id_out) { set_error("comm: NULL id buffer"); return 1; }
  NcclId id;
  UVGB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

extern "C" uvg_b200_comm *uvg_b200_comm_create(const void *id, int rank, int world)
{
  if (ensure_init() || need_nccl()) return nullptr;
  if (!id || world < 1 || rank < 0 || rank >= world) { set_error("comm: bad arguments (rank %d of %d)", rank, world); return nullptr; }
  NcclId nid;
  memcpy(&nid, id, sizeof(nid));
  uvg_b200_comm *c = new (std::nothrow) uvg_b200_comm();
  if (!c) { set_error("comm: out of memory"); return nullptr; }
  c->rank = rank; c->world = world; c->comm = nullptr;
  const int r = g_nccl.CommInitRank(&c->comm, world, nid, rank);
  if (r != 0) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); delete c; return nullptr; }
  return c;
}

extern "C" void uvg_b200_comm_destroy(uvg_b200_comm *c)
{
  if (!c) return;
  if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
  delete c;
}

extern "C" int uvg_b200_comm_nccl_version(void)
{
  if (need_nccl() || !g_nccl.GetVersion) return 0;
  int v = 0;
  g_nccl.GetVersion(&v);
  return v;
}

// In place on `plane` (device, `stride_bytes` per pixel row, `height` rows): rank r has written the pixel rows of its
// band [lo_r*64, min(hi_r*64, height)); on return (stream order) every rank holds all bands.
extern "C" int uvg_b200_refpic_exchange(uvg_b200_comm *c, void *plane, size_t stride_bytes, int height, void *stream)
{
  if (!c || !plane) { set_error("refpic_exchange: NULL argument"); return 1; }
  if (c->world == 1) return 0;
  const int ctu_rows = (height + 63) / 64;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UVGB_NCCL(g_nccl.GroupStart());
  for (int r = 0; r < c->world; ++r) {
    int lo, hi;
    uvg_b200_band_rows(r, c->world, ctu_rows, &lo, &hi);
    const int y0 = lo * 64, y1 = hi * 64 < height ? hi * 64 : height;
    if (y1 <= y0) continue;
    uint8_t *p = static_cast<uint8_t *>(plane) + (size_t)y0 * stride_bytes;
    const int rc = g_nccl.Broadcast(p, p, (size_t)(y1 - y0) * stride_bytes, 1 /* ncclUint8 */, r, c->comm, st);
    if (rc != 0) { g_nccl.GroupEnd(); set_error("ncclBroadcast failed: %s", g_nccl.GetErrorString(rc)); return 1; }
  }
  UVGB_NCCL(g_nccl.GroupEnd());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
