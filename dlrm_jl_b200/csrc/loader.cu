// Device-side batch marshalling for the Criteo "DAC" record stream.
//
// SURVEY.md section 8(f) row 2.  The reference's `load!` (DLRM.jl src/data/criteo.jl:284-310) is a
// threaded CPU loop that copies, per record, the label, 13 continuous floats into dense[13 x B] and
// 26 categorical UInt32 into sparse[B x 26].  Here the raw 160-byte records of a batch
// (struct DACRecord, src/data/criteo.jl:91-95: Int32 label, 13 x Float32, 26 x UInt32, packed) are
// copied to the device as ONE contiguous block and unpacked there: labels [B] (as Float32, what
// bce_loss consumes), dense [B][13], sparse [26][B] -- the table-major index layout
// dlrmb_embedding_fwd takes.  A tile of records is staged through shared memory so both the record
// reads and the transposed index writes are coalesced.
#include "common.cuh"

namespace dlrmb {

constexpr int kRecWords = 40;   // 160 bytes
constexpr int kRecTile = 128;   // records per CTA

__global__ void __launch_bounds__(256)
dac_unpack_kernel(const uint32_t* __restrict__ rec, int B, float* __restrict__ labels,
                  float* __restrict__ dense, uint32_t* __restrict__ sparse) {
    __shared__ uint32_t tile[kRecTile * kRecWords + 1];
    const int b0 = blockIdx.x * kRecTile;
    const int nb = min(kRecTile, B - b0);
    const uint32_t* src = rec + (size_t)b0 * kRecWords;
    for (int i = threadIdx.x; i < nb * kRecWords; i += blockDim.x) tile[i] = src[i];
    __syncthreads();
    // label: Int32 -> Float32
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        labels[b0 + i] = (float)(int32_t)tile[i * kRecWords];
    // dense [B][13]: contiguous per record
    for (int i = threadIdx.x; i < nb * 13; i += blockDim.x) {
        const int b = i / 13, j = i - b * 13;
        dense[(size_t)(b0 + b) * 13 + j] = __uint_as_float(tile[b * kRecWords + 1 + j]);
    }
    // sparse [26][B]: b fastest, so each table's row is written contiguously
    for (int i = threadIdx.x; i < nb * 26; i += blockDim.x) {
        const int j = i / nb, b = i - j * nb;
        sparse[(size_t)j * B + b0 + b] = tile[b * kRecWords + 14 + j];
    }
}

int launch_dac_unpack(const void* rec, int B, float* labels, float* dense, uint32_t* sparse, cudaStream_t s) {
    const int blocks = (B + kRecTile - 1) / kRecTile;
    dac_unpack_kernel<<<blocks, 256, 0, s>>>(static_cast<const uint32_t*>(rec), B, labels, dense, sparse);
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

}  // namespace dlrmb

extern "C" int32_t dlrmb_dac_unpack(int32_t device, const void* records, int32_t B, float* labels,
                                    float* dense, uint32_t* sparse, dlrmb_stream stream) {
    using namespace dlrmb;
    DLRMB_REQUIRE(B > 0, "B must be positive (got %d)", B);
    DLRMB_REQUIRE(records && labels && dense && sparse, "null buffer");
    DLRMB_REQUIRE((reinterpret_cast<uintptr_t>(records) & 3) == 0, "records must be 4-byte aligned");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) DLRMB_CUDA(cudaSetDevice(device));
    int rc = launch_dac_unpack(records, B, labels, dense, sparse, (cudaStream_t)stream);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}
