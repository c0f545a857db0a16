// Peer-memory exchange over NVLink / NVSwitch for the K-sharded queue: ONE kernel pushes this rank's rows
// straight into every peer's symmetric buffer (remote stores), waits for the peers' rows and hands them to the next
// kernel -- the all-gather of the queries (fused with their bf16 cast), the all-gather of the keys and the all-to-all
// of the partial InfoNCE records, without a library collective on the critical path.
//
// Every rank owns an identical "symmetric" allocation (torch.distributed._symmetric_memory) and knows the base
// address of every peer's copy.  Layout of that allocation, same on all ranks:
//   [ctrl_off, +sizeof(PeerCtrl))  control block (zero-initialised once)
//   [data_off + (2 * channel + parity) * region_bytes, ...)   receive regions, slot r = the rows pushed by rank r
// Low-latency protocol (flag in band, no fence / signal round trip): every 4-byte payload word travels as an 8-byte
// (word, epoch tag) pair, two pairs per 16-byte store; the receiver polls the pairs themselves until both tags equal
// the current epoch and writes the payload to `out` -- so a message costs one NVLink store latency plus the poll.
//   e = epoch[c] + 1 (device-resident, so a captured CUDA graph replays correctly), parity = e & 1
//   push:   rank r expands its rows for peer p into p's region[c][parity] slot r     (plain remote 16-byte stores)
//   poll:   region[c][parity] -> `out` (a fixed local buffer: the consumer's pointer never changes between replays)
//   last CTA: epoch[c] = e
// A stale pair can never match: its tag is an older epoch (the regions start zeroed, epochs start at 1).  Two
// parities make the scheme safe without a barrier: a rank can start writing epoch e + 2 into a region only after it
// has received every peer's epoch e + 1 rows, which a peer pushes after it has finished reading epoch e.
// Every CTA polls remote data, so the grid must be co-resident: it is capped well below one CTA per SM.
#include "common.cuh"

namespace moma {

namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ int g_peer_error = 0;

// pairs_per_rank: 8-byte units of OUTPUT payload per rank slot (two 4-byte words each).  CAST: the source is fp32
// (16 bytes of source per output pair).  Each pair of payload words occupies one 16-byte cell (w0, tag, w1, tag).
template <bool CAST>
__global__ void __launch_bounds__(kPeerThreads)
peer_exchange_kernel(const void* __restrict__ src, long long src_peer_stride_bytes, long long pairs_per_rank,
                     const unsigned long long* __restrict__ peer_bases, long long ctrl_off, long long data_off,
                     long long region_bytes, int rank, int world, int channel, void* __restrict__ out) {
    pdl_wait();
    const unsigned long long my_base = peer_bases[rank];
    PeerCtrl* ctrl = reinterpret_cast<PeerCtrl*>(my_base + ctrl_off);
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch[channel]) + 1ull;
    const uint32_t tag = (uint32_t)e;
    const long long region = data_off + (2ll * channel + (long long)(e & 1ull)) * region_bytes;
    const long long total = pairs_per_rank * world;
    const long long stride = (long long)gridDim.x * kPeerThreads;

    // ---- push: pair v of this rank's rows for peer p -> cell v of slot `rank` in p's region.  The pair index is the
    // outer loop (no integer division per element); for an all-gather (stride 0) the source is read once for all peers.
    const long long slot = region + (long long)rank * pairs_per_rank * 16;
    for (long long v = (long long)blockIdx.x * kPeerThreads + threadIdx.x; v < pairs_per_rank; v += stride) {
        uint2 w = make_uint2(0u, 0u);
        for (int p = 0; p < world; ++p) {
            if (p == 0 || src_peer_stride_bytes != 0) {
                const char* s = static_cast<const char*>(src) + (long long)p * src_peer_stride_bytes;
                if (CAST) {
                    const float4 f = *(reinterpret_cast<const float4*>(s) + v);
                    w = make_uint2(pack2(f.x, f.y), pack2(f.z, f.w));
                } else {
                    w = *(reinterpret_cast<const uint2*>(s) + v);
                }
            }
            *reinterpret_cast<uint4*>(peer_bases[p] + slot + v * 16) = make_uint4(w.x, tag, w.y, tag);
        }
    }
    pdl_launch_dependents();

    // ---- poll the cells the peers are writing and hand the payload to the consumer
    const uint4* in = reinterpret_cast<const uint4*>(my_base + region);
    uint2* o2 = static_cast<uint2*>(out);
    const long long t0 = clock64();
    for (long long i = (long long)blockIdx.x * kPeerThreads + threadIdx.x; i < total; i += stride) {
        uint4 c = ld_volatile16(in + i);
        unsigned ns = 32;
        while (c.y != tag || c.w != tag) {
            __nanosleep(ns);                                  // back off: the step's other branches share these SMs
            if (ns < 256) ns *= 2;
            if (clock64() - t0 > kPeerTimeoutClk) { atomicExch(&g_peer_error, 1 + channel); __trap(); }
            c = ld_volatile16(in + i);
        }
        o2[i] = make_uint2(c.x, c.z);
    }

    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&ctrl->ticket_done[channel], 1u) == gridDim.x - 1u) {      // every CTA has read epoch[channel]
            ctrl->ticket_done[channel] = 0u;
            *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch[channel]) = e;
        }
    }
}

}  // namespace

}  // namespace moma

using namespace moma;

extern "C" __attribute__((visibility("default"))) size_t moma_peer_ctrl_bytes(void) { return (sizeof(PeerCtrl) + 255) / 256 * 256; }

// src: this rank's rows; peer p receives the bytes_per_rank bytes (of OUTPUT) starting at src + p * src_peer_stride_bytes
// (stride 0 = all-gather, stride = one block = all-to-all).  cast_f32_to_bf16: src is fp32, the pushed rows are bf16
// (bytes_per_rank counts the bf16 bytes).  out: [world, bytes_per_rank] local buffer, slot s = the rows from rank s.
extern "C" __attribute__((visibility("default"))) int moma_peer_exchange(
    const void* src, int64_t src_peer_stride_bytes, int64_t bytes_per_rank, int cast_f32_to_bf16,
    const uint64_t* peer_bases_dev, int64_t ctrl_off, int64_t data_off, int64_t region_bytes, int rank, int world,
    int channel, void* out, moma_stream_t stream) {
    MOMA_REQUIRE(src && peer_bases_dev && out, MOMA_ERR_INVALID, "peer_exchange: null pointer");
    MOMA_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, MOMA_ERR_INVALID,
                 "peer_exchange: bad rank/world %d/%d (max world %d)", rank, world, kPeerMaxWorld);
    MOMA_REQUIRE(channel >= 0 && channel < kPeerChannels, MOMA_ERR_INVALID, "peer_exchange: channel %d out of range", channel);
    MOMA_REQUIRE(bytes_per_rank > 0 && bytes_per_rank % 16 == 0 && src_peer_stride_bytes % 16 == 0 && src_peer_stride_bytes >= 0,
                 MOMA_ERR_ALIGN, "peer_exchange: sizes must be multiples of 16 bytes");
    MOMA_REQUIRE(2 * bytes_per_rank * world <= region_bytes && region_bytes % 16 == 0 && data_off % 16 == 0 && ctrl_off % 16 == 0,
                 MOMA_ERR_WORKSPACE, "peer_exchange: region too small (needs 2 x world x bytes_per_rank) or unaligned");
    MOMA_REQUIRE(aligned16(src) && aligned16(out), MOMA_ERR_ALIGN, "peer_exchange: src/out must be 16-byte aligned");
    const long long pairs = bytes_per_rank / 8;
    long long ctas = (pairs * world + kPeerThreads * 2 - 1) / (kPeerThreads * 2);
    const long long cap = sm_count() / 4;                      // co-resident by construction
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    cudaStream_t st = as_stream(stream);
    const unsigned long long* bases = reinterpret_cast<const unsigned long long*>(peer_bases_dev);
    if (cast_f32_to_bf16)
        launch_pdl(peer_exchange_kernel<true>, dim3((unsigned)ctas), dim3(kPeerThreads), 0, st, src, (long long)src_peer_stride_bytes,
                   pairs, bases, (long long)ctrl_off, (long long)data_off, (long long)region_bytes, rank, world, channel, out);
    else
        launch_pdl(peer_exchange_kernel<false>, dim3((unsigned)ctas), dim3(kPeerThreads), 0, st, src, (long long)src_peer_stride_bytes,
                   pairs, bases, (long long)ctrl_off, (long long)data_off, (long long)region_bytes, rank, world, channel, out);
    MOMA_CUDA_LAUNCH_CHECK("peer_exchange");
    note_launches(1);
    return MOMA_OK;
}
