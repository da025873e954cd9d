// cusparse_check.cu -- TEST INFRASTRUCTURE: an independent GPU opinion on C = A*B from cuSPARSE's SpGEMM, in FP64.
// The reference compares its result with cusparseSpGEMM too (src/external/cusparse/spgemm_cusparse.h:99-325) but creates
// its descriptors with CUDA_R_32F on double data (:39,165-173) and has the value comparison commented out (:282); this is
// the same call sequence with CUDA_R_64F. Never linked into the product: tests/test_cusparse_harness.py builds it with
// nvcc -lcusparse, feeds it two CSR files and compares what it writes with libtilespgemm_b200's result.
//   usage: cusparse_check A.csr B.csr C.csr      (file: int32 m, n, nnz | rowptr[m+1] | colidx[nnz] | val f64[nnz])
#include <cuda_runtime.h>
#include <cusparse.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CHECK(x) do { if ((x) != 0) { fprintf(stderr, "%s failed at line %d\n", #x, __LINE__); return 2; } } while (0)

struct Csr { int m, n, nnz; std::vector<int> rp, ci; std::vector<double> v; };

static int read_csr(const char *path, Csr &a)
{
    FILE *f = fopen(path, "rb");
    if (!f) return 1;
    int h[3];
    if (fread(h, 4, 3, f) != 3) return 1;
    a.m = h[0]; a.n = h[1]; a.nnz = h[2];
    a.rp.resize(a.m + 1); a.ci.resize(a.nnz > 0 ? a.nnz : 1); a.v.resize(a.nnz > 0 ? a.nnz : 1);
    int ok = fread(a.rp.data(), 4, a.m + 1, f) == (size_t)a.m + 1 && fread(a.ci.data(), 4, a.nnz, f) == (size_t)a.nnz &&
             fread(a.v.data(), 8, a.nnz, f) == (size_t)a.nnz;
    fclose(f);
    return !ok;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: cusparse_check A.csr B.csr C.csr\n"); return 1; }
    Csr A, B;
    if (read_csr(argv[1], A) || read_csr(argv[2], B) || A.n != B.m) { fprintf(stderr, "bad input\n"); return 1; }
    int *dArp, *dAci, *dBrp, *dBci, *dCrp, *dCci = nullptr;
    double *dAv, *dBv, *dCv = nullptr;
    CHECK(cudaMalloc(&dArp, (A.m + 1) * 4)); CHECK(cudaMalloc(&dAci, A.ci.size() * 4)); CHECK(cudaMalloc(&dAv, A.v.size() * 8));
    CHECK(cudaMalloc(&dBrp, (B.m + 1) * 4)); CHECK(cudaMalloc(&dBci, B.ci.size() * 4)); CHECK(cudaMalloc(&dBv, B.v.size() * 8));
    CHECK(cudaMalloc(&dCrp, (A.m + 1) * 4));
    CHECK(cudaMemcpy(dArp, A.rp.data(), (A.m + 1) * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dAci, A.ci.data(), A.ci.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dAv, A.v.data(), A.v.size() * 8, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dBrp, B.rp.data(), (B.m + 1) * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dBci, B.ci.data(), B.ci.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dBv, B.v.data(), B.v.size() * 8, cudaMemcpyHostToDevice));

    cusparseHandle_t h;
    cusparseSpMatDescr_t mA, mB, mC;
    CHECK(cusparseCreate(&h));
    CHECK(cusparseCreateCsr(&mA, A.m, A.n, A.nnz, dArp, dAci, dAv, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
    CHECK(cusparseCreateCsr(&mB, B.m, B.n, B.nnz, dBrp, dBci, dBv, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
    CHECK(cusparseCreateCsr(&mC, A.m, B.n, 0, dCrp, nullptr, nullptr, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
    const double alpha = 1.0, beta = 0.0;
    cusparseSpGEMMDescr_t desc;
    CHECK(cusparseSpGEMM_createDescr(&desc));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    size_t b1 = 0, b2 = 0;
    void *w1 = nullptr, *w2 = nullptr;
    const cusparseOperation_t N = CUSPARSE_OPERATION_NON_TRANSPOSE;
    CHECK(cusparseSpGEMM_workEstimation(h, N, N, &alpha, mA, mB, &beta, mC, CUDA_R_64F, CUSPARSE_SPGEMM_DEFAULT, desc, &b1, nullptr));
    CHECK(cudaMalloc(&w1, b1 ? b1 : 1));
    CHECK(cusparseSpGEMM_workEstimation(h, N, N, &alpha, mA, mB, &beta, mC, CUDA_R_64F, CUSPARSE_SPGEMM_DEFAULT, desc, &b1, w1));
    CHECK(cusparseSpGEMM_compute(h, N, N, &alpha, mA, mB, &beta, mC, CUDA_R_64F, CUSPARSE_SPGEMM_DEFAULT, desc, &b2, nullptr));
    CHECK(cudaMalloc(&w2, b2 ? b2 : 1));
    CHECK(cusparseSpGEMM_compute(h, N, N, &alpha, mA, mB, &beta, mC, CUDA_R_64F, CUSPARSE_SPGEMM_DEFAULT, desc, &b2, w2));
    int64_t cm, cn, cnnz;
    CHECK(cusparseSpMatGetSize(mC, &cm, &cn, &cnnz));
    CHECK(cudaMalloc(&dCci, (cnnz ? cnnz : 1) * 4)); CHECK(cudaMalloc(&dCv, (cnnz ? cnnz : 1) * 8));
    CHECK(cusparseCsrSetPointers(mC, dCrp, dCci, dCv));
    CHECK(cusparseSpGEMM_copy(h, N, N, &alpha, mA, mB, &beta, mC, CUDA_R_64F, CUSPARSE_SPGEMM_DEFAULT, desc));
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<int> crp(A.m + 1), cci(cnnz ? cnnz : 1);
    std::vector<double> cv(cnnz ? cnnz : 1);
    CHECK(cudaMemcpy(crp.data(), dCrp, (A.m + 1) * 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(cci.data(), dCci, (size_t)cnnz * 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(cv.data(), dCv, (size_t)cnnz * 8, cudaMemcpyDeviceToHost));
    FILE *f = fopen(argv[3], "wb");
    if (!f) return 1;
    int hdr[3] = {A.m, B.n, (int)cnnz};
    fwrite(hdr, 4, 3, f); fwrite(crp.data(), 4, A.m + 1, f); fwrite(cci.data(), 4, cnnz, f); fwrite(cv.data(), 8, cnnz, f);
    fclose(f);
    printf("cusparse SpGEMM (FP64): nnzC = %lld, %.3f ms (work estimation + compute + copy, first call)\n", (long long)cnnz, ms);
    return 0;
}
