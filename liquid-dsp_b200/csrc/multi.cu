// multi.cu -- C-ABI multi-GPU driver for the time-sharded channelizer: one process
// per GPU, rank r owns the frames [first_frame, first_frame + n_frames) of ONE long stream, every segment
// starts on an even frame (the reference's flag = 0, firpfbch2.c:57).  One exchange step per pass: the last
// H = 2mM input samples of rank r's segment go to rank r + 1 (ncclSend / ncclRecv in one group, NVLink),
// straight into the head-room in front of its input, so the kernels see one contiguous (halo | segment)
// array.  The interior frames (those at least 4m frames from the segment start) do not see the halo: they
// are launched first on the caller's stream, the exchange and the 4m halo-dependent frames run beside them
// on the communicator's side stream, and the caller's stream joins the side stream at the end.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy a host process -- e.g. PyTorch -- has
// already loaded, else the system one), so libliquid_cuda.so itself links against nothing but the CUDA
// runtime and single-GPU programs never touch NCCL.
#include <dlfcn.h>

#include <cstdint>

#include "lc_common.cuh"
#include "../../include/liquid_cuda.h"

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: nccl.h:37-41, 285)
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm *NcclComm;
constexpr int kNcclFloat = 7;

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

NcclApi &nccl()
{
    static NcclApi api;
    if (api.handle) return api;
    const char *names[] = {getenv("LIQUID_CUDA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n) continue;
        // RTLD_NOLOAD first: reuse the copy the process already carries (PyTorch bundles its own)
        api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
        if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        fprintf(stderr, "liquid_cuda: cannot load NCCL (libnccl.so.2): %s; the multi-GPU driver needs it "
                        "(set LIQUID_CUDA_NCCL_LIB to its path)\n", dlerror());
        abort();
    }
    auto sym = [&](const char *s) {
        void *p = dlsym(api.handle, s);
        if (!p) {
            fprintf(stderr, "liquid_cuda: NCCL symbol %s not found\n", s);
            abort();
        }
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}

#define LC_NCCL_CHECK(expr)                                                                                   \
    do {                                                                                                      \
        int lc_r__ = (expr);                                                                                  \
        if (lc_r__ != 0) {                                                                                    \
            fprintf(stderr, "liquid_cuda: NCCL error %d (%s) at %s:%d: %s\n", lc_r__,                         \
                    nccl().GetErrorString(lc_r__), __FILE__, __LINE__, #expr);                                \
            abort();                                                                                          \
        }                                                                                                     \
    } while (0)

}  // namespace

struct liquid_cuda_comm_s {
    int rank = 0, world = 1;
    NcclComm comm = nullptr;
    cudaStream_t side = nullptr;       // the exchange and the halo-dependent frames
    cudaEvent_t ev_in = nullptr;       // caller's stream: the segment's input is in place
    cudaEvent_t ev_side = nullptr;     // side stream: halo frames written
    int saved_reserved_sms = 0;
};

extern "C" {

int liquid_cuda_comm_unique_id_bytes(void) { return (int)sizeof(NcclUniqueId); }

void liquid_cuda_comm_get_unique_id(void *_id)
{
    lc::ensure_device();
    NcclUniqueId id;
    LC_NCCL_CHECK(nccl().GetUniqueId(&id));
    memcpy(_id, &id, sizeof id);
}

liquid_cuda_comm liquid_cuda_comm_create(int _rank, int _world, const void *_id)
{
    if (_world < 1 || _rank < 0 || _rank >= _world) lc::fatal("liquid_cuda_comm_create", "rank must be in [0, world)");
    lc::ensure_device();
    liquid_cuda_comm c = new liquid_cuda_comm_s();
    c->rank = _rank;
    c->world = _world;
    LC_CUDA_CHECK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    LC_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
    LC_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
    c->saved_reserved_sms = liquid_cuda_get_reserved_sms();
    if (_world > 1) {
        if (!_id) lc::fatal("liquid_cuda_comm_create", "a unique id is required when world > 1");
        // The halo is 112-576 KB: one NCCL channel (one CTA per rank) moves it.  With the default channel count
        // the send/recv kernel is a grid of CTAs that must be co-resident with their peers'; beside a persistent
        // kernel that owns almost every SM they are not, and ranks stall each other for a whole kernel duration.
        setenv("NCCL_MAX_P2P_NCHANNELS", "1", 0);
        setenv("NCCL_MIN_P2P_NCHANNELS", "1", 0);
        NcclUniqueId id;
        memcpy(&id, _id, sizeof id);
        LC_NCCL_CHECK(nccl().CommInitRank(&c->comm, _world, id, _rank));
        // the persistent one-CTA-per-SM kernels leave two SMs free: one for the NCCL CTA, one for the halo launch
        if (c->saved_reserved_sms < 2) liquid_cuda_set_reserved_sms(2);
    }
    return c;
}

void liquid_cuda_comm_destroy(liquid_cuda_comm _c)
{
    if (!_c) return;
    cudaStreamSynchronize(_c->side);
    if (_c->comm) {
        nccl().CommDestroy(_c->comm);
        liquid_cuda_set_reserved_sms(_c->saved_reserved_sms);
    }
    cudaEventDestroy(_c->ev_in);
    cudaEventDestroy(_c->ev_side);
    cudaStreamDestroy(_c->side);
    delete _c;
}

int liquid_cuda_comm_rank(liquid_cuda_comm _c) { return _c->rank; }
int liquid_cuda_comm_world(liquid_cuda_comm _c) { return _c->world; }

void liquid_cuda_comm_exchange_halo_dev(liquid_cuda_comm _c, void *_d_head, const void *_d_tail,
                                        unsigned long long _n_samples, void *_stream)
{
    if (_c->world == 1 || _n_samples == 0) return;
    cudaStream_t st = lc::as_stream(_stream);
    LC_NCCL_CHECK(nccl().GroupStart());
    if (_c->rank + 1 < _c->world) LC_NCCL_CHECK(nccl().Send(_d_tail, 2 * _n_samples, kNcclFloat, _c->rank + 1, _c->comm, st));
    if (_c->rank > 0) LC_NCCL_CHECK(nccl().Recv(_d_head, 2 * _n_samples, kNcclFloat, _c->rank - 1, _c->comm, st));
    LC_NCCL_CHECK(nccl().GroupEnd());
}

void firpfbch2_crcf_execute_sharded_dev(firpfbch2_crcf _q, liquid_cuda_comm _c, void *_d_x_with_headroom,
                                        unsigned long long _first_frame, unsigned long long _n_frames,
                                        void *_d_y, void *_stream)
{
    if (_first_frame & 1) lc::fatal("firpfbch2_crcf_execute_sharded_dev", "a segment must start on an even frame");
    cudaStream_t st = lc::as_stream(_stream);
    const unsigned int H = firpfbch2_crcf_get_state_len(_q);
    const unsigned int M2 = firpfbch2_crcf_get_num_channels(_q) / 2, M = 2 * M2;
    float2 *d_halo = static_cast<float2 *>(_d_x_with_headroom);
    float2 *d_x = d_halo + H;
    float2 *d_y = static_cast<float2 *>(_d_y);
    if (_c->world == 1) {
        firpfbch2_crcf_execute_segment_dev(_q, d_halo, d_x, _first_frame, _n_frames, d_y, st);
        return;
    }
    unsigned long long halo_frames = H / M2;               // 4m frames read the head-room
    if (halo_frames > _n_frames) halo_frames = _n_frames;
    if ((unsigned long long)H > _n_frames * M2)
        lc::fatal("firpfbch2_crcf_execute_sharded_dev", "a segment must hold at least the 2mM samples its right neighbour needs");
    // side stream: exchange, then the frames that read the head-room
    LC_CUDA_CHECK(cudaEventRecord(_c->ev_in, st));
    LC_CUDA_CHECK(cudaStreamWaitEvent(_c->side, _c->ev_in, 0));
    liquid_cuda_comm_exchange_halo_dev(_c, d_halo, d_x + _n_frames * M2 - H, H, _c->side);
    // caller's stream: the interior frames, while the halo is in flight
    if (_n_frames > halo_frames)
        firpfbch2_crcf_execute_segment_dev(_q, d_x + halo_frames * M2 - H, d_x + halo_frames * M2, _first_frame + halo_frames,
                                           _n_frames - halo_frames, d_y + halo_frames * M, st);
    firpfbch2_crcf_execute_segment_dev(_q, d_halo, d_x, _first_frame, halo_frames, d_y, _c->side);
    LC_CUDA_CHECK(cudaEventRecord(_c->ev_side, _c->side));
    LC_CUDA_CHECK(cudaStreamWaitEvent(st, _c->ev_side, 0));
}

}  // extern "C"
