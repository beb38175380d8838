// pipeline.cu — end-to-end regrid of host-resident fields: the work behind one
// RegridFilter.forward(FieldList) call (reference filters/fields/regrid.py:174-208, whose
// loop applies the matrix field by field on the CPU).
//
// The FieldList hands over one host array per field (field-major).  Fields are processed in
// chunks; per chunk
//     copy stream   : H2D of the chunk's fields into a field-major staging buffer
//     compute stream: pack (transpose to point-major) -> SpMM -> unpack (transpose back)
//     drain stream  : D2H of the regridded fields
// with kBuffers staging buffers in each direction so the three stages of consecutive chunks
// overlap (PCIe is full duplex).  Events carry the dependencies; the host only blocks once,
// at the end.
#include <algorithm>
#include <vector>

#include "common.cuh"

struct at_csr;  // defined in spmm.cu; only accessed through the C-ABI here

namespace {
constexpr int kBuffers = 3;
}

struct at_pipeline {
    const at_csr_t* csr = nullptr;
    int64_t n_src = 0, n_tgt = 0;
    int chunk = 0;
    float* d_in[kBuffers] = {nullptr, nullptr, nullptr};   // [chunk, n_src] field-major
    float* d_out[kBuffers] = {nullptr, nullptr, nullptr};  // [chunk, n_tgt] field-major
    float* d_x = nullptr;                                  // [n_src, chunk] point-major
    float* d_y = nullptr;                                  // [n_tgt, chunk] point-major
    cudaStream_t s_in = nullptr, s_compute = nullptr, s_out = nullptr;
    cudaEvent_t h2d_done[kBuffers] = {nullptr, nullptr, nullptr};
    cudaEvent_t packed[kBuffers] = {nullptr, nullptr, nullptr};
    cudaEvent_t computed[kBuffers] = {nullptr, nullptr, nullptr};
    cudaEvent_t drained[kBuffers] = {nullptr, nullptr, nullptr};
};

using namespace at;

extern "C" int at_pipeline_destroy(at_pipeline_t* p) {
    if (p == nullptr) return AT_OK;
    for (int b = 0; b < kBuffers; ++b) {
        cudaFree(p->d_in[b]);
        cudaFree(p->d_out[b]);
        if (p->h2d_done[b]) cudaEventDestroy(p->h2d_done[b]);
        if (p->packed[b]) cudaEventDestroy(p->packed[b]);
        if (p->computed[b]) cudaEventDestroy(p->computed[b]);
        if (p->drained[b]) cudaEventDestroy(p->drained[b]);
    }
    cudaFree(p->d_x);
    cudaFree(p->d_y);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_compute) cudaStreamDestroy(p->s_compute);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
    return AT_OK;
}

extern "C" int at_pipeline_create(const at_csr_t* csr, int32_t chunk_fields, at_pipeline_t** out) {
    AT_REQUIRE(out != nullptr, "at_pipeline_create: out is null");
    *out = nullptr;
    AT_REQUIRE(csr != nullptr, "at_pipeline_create: null matrix");
    AT_REQUIRE(chunk_fields >= 4 && chunk_fields % 4 == 0 && chunk_fields <= 4096,
               "at_pipeline_create: chunk_fields must be a multiple of 4 in [4, 4096]");
    int64_t n_rows, n_cols, nnz;
    int uniform, dtype;
    int rc = at_csr_info(csr, &n_rows, &n_cols, &nnz, &uniform, &dtype);
    if (rc != AT_OK) return rc;
    AT_REQUIRE(dtype == AT_F32, "at_pipeline_create: float32 matrices only");

    at_pipeline* p = new at_pipeline();
    p->csr = csr;
    p->n_src = n_cols;
    p->n_tgt = n_rows;
    p->chunk = chunk_fields;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
        return e == cudaSuccess;
    };
    const size_t in_bytes = static_cast<size_t>(n_cols) * chunk_fields * 4;
    const size_t out_bytes = static_cast<size_t>(n_rows) * chunk_fields * 4;
    for (int b = 0; b < kBuffers && e == cudaSuccess; ++b) {
        ok(cudaMalloc(&p->d_in[b], std::max<size_t>(in_bytes, 16)));
        ok(cudaMalloc(&p->d_out[b], std::max<size_t>(out_bytes, 16)));
        ok(cudaEventCreateWithFlags(&p->h2d_done[b], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&p->packed[b], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&p->computed[b], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&p->drained[b], cudaEventDisableTiming));
    }
    ok(cudaMalloc(&p->d_x, std::max<size_t>(in_bytes, 16)));
    ok(cudaMalloc(&p->d_y, std::max<size_t>(out_bytes, 16)));
    ok(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    if (e != cudaSuccess) {
        at_pipeline_destroy(p);
        return set_error(e == cudaErrorMemoryAllocation ? AT_ERR_NOMEM : AT_ERR_CUDA,
                         "at_pipeline_create: %s", cudaGetErrorString(e));
    }
    *out = p;
    return AT_OK;
}

extern "C" int at_pipeline_regrid(at_pipeline_t* p, const float* const* fields_in, float* const* fields_out,
                                  int64_t n_fields) {
    AT_REQUIRE(p != nullptr && fields_in != nullptr && fields_out != nullptr, "at_pipeline_regrid: null argument");
    AT_REQUIRE(n_fields >= 0, "at_pipeline_regrid: negative field count");
    for (int64_t f = 0; f < n_fields; ++f)
        AT_REQUIRE(fields_in[f] != nullptr && fields_out[f] != nullptr, "at_pipeline_regrid: field %lld is null",
                   (long long)f);
    const int64_t n_chunks = (n_fields + p->chunk - 1) / p->chunk;
    const size_t src_bytes = static_cast<size_t>(p->n_src) * 4, tgt_bytes = static_cast<size_t>(p->n_tgt) * 4;
    int rc = AT_OK;
    for (int64_t c = 0; c < n_chunks && rc == AT_OK; ++c) {
        const int b = static_cast<int>(c % kBuffers);
        const int64_t f0 = c * p->chunk;
        const int nf = static_cast<int>(std::min<int64_t>(p->chunk, n_fields - f0));
        // stage in: the buffer is free once the pack of chunk c - kBuffers has read it
        if (c >= kBuffers) AT_CUDA_TRY(cudaStreamWaitEvent(p->s_in, p->packed[b], 0));
        // fields that are adjacent in host memory (one [F, n] array) go in one copy
        for (int f = 0; f < nf;) {
            int g = f + 1;
            while (g < nf && fields_in[f0 + g] == fields_in[f0 + g - 1] + p->n_src) ++g;
            AT_CUDA_TRY(cudaMemcpyAsync(p->d_in[b] + static_cast<size_t>(f) * p->n_src, fields_in[f0 + f],
                                        src_bytes * static_cast<size_t>(g - f), cudaMemcpyHostToDevice, p->s_in));
            f = g;
        }
        AT_CUDA_TRY(cudaEventRecord(p->h2d_done[b], p->s_in));

        // compute
        AT_CUDA_TRY(cudaStreamWaitEvent(p->s_compute, p->h2d_done[b], 0));
        rc = at_transpose(p->d_in[b], nf, p->n_src, p->n_src, p->d_x, p->chunk, 4, p->s_compute);
        if (rc != AT_OK) break;
        AT_CUDA_TRY(cudaEventRecord(p->packed[b], p->s_compute));
        rc = at_spmm(p->csr, p->d_x, AT_F32, p->chunk, p->d_y, AT_F32, p->chunk, nf, 0, p->s_compute);
        if (rc != AT_OK) break;
        // the output staging buffer is free once chunk c - kBuffers has been drained
        if (c >= kBuffers) AT_CUDA_TRY(cudaStreamWaitEvent(p->s_compute, p->drained[b], 0));
        rc = at_transpose(p->d_y, p->n_tgt, nf, p->chunk, p->d_out[b], p->n_tgt, 4, p->s_compute);
        if (rc != AT_OK) break;
        AT_CUDA_TRY(cudaEventRecord(p->computed[b], p->s_compute));

        // drain
        AT_CUDA_TRY(cudaStreamWaitEvent(p->s_out, p->computed[b], 0));
        for (int f = 0; f < nf;) {
            int g = f + 1;
            while (g < nf && fields_out[f0 + g] == fields_out[f0 + g - 1] + p->n_tgt) ++g;
            AT_CUDA_TRY(cudaMemcpyAsync(fields_out[f0 + f], p->d_out[b] + static_cast<size_t>(f) * p->n_tgt,
                                        tgt_bytes * static_cast<size_t>(g - f), cudaMemcpyDeviceToHost, p->s_out));
            f = g;
        }
        AT_CUDA_TRY(cudaEventRecord(p->drained[b], p->s_out));
    }
    cudaError_t e1 = cudaStreamSynchronize(p->s_in);
    cudaError_t e2 = cudaStreamSynchronize(p->s_compute);
    cudaError_t e3 = cudaStreamSynchronize(p->s_out);
    if (rc != AT_OK) return rc;
    const cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
    if (e != cudaSuccess) return set_error(AT_ERR_CUDA, "at_pipeline_regrid: %s", cudaGetErrorString(e));
    return AT_OK;
}
