// sddmm.hpp -- entry points of the reference's include/sddmm.hpp:8-21, include/sddmmKernel.cuh:19-51 and
// include/host.hpp (checker), forwarding to libsddmm_b200.
#pragma once
#include "BSMR.hpp"
#include "Logger.hpp"
#include "Matrix.hpp"
#include "Options.hpp"

void sddmm(const Options& options, const Matrix<float>& matrixA, const Matrix<float>& matrixB,
           sparseMatrix::CSR<float>& matrixP, Logger& logger);
void sddmm_testMode(const Options& options, sparseMatrix::CSR<float>& matrixP);
// Multi-GPU form of sddmm() (no reference counterpart: the reference is single-GPU).  Called by EVERY rank's
// process with the same S, A, B: rank 0 reorders, sddmm_mgpu_shard hands every rank its nnz-balanced row-panel
// range, B is replicated once over NCCL, the timed passes run without any collective, sddmm_mgpu_gather leaves
// the whole P on every rank.  Options: -r rank -w world -u id-file (rank 0 writes the NCCL id there).
bool sddmm_multiGpu(const Options& options, const Matrix<float>& matrixA, const Matrix<float>& matrixB,
                    sparseMatrix::CSR<float>& matrixP, Logger& logger);
void sddmm_gpu(const Matrix<float>& matrixA, const Matrix<float>& matrixB, const RPHM& rphm,
               sparseMatrix::CSR<float>& matrixP, Logger& logger);
// raw device-pointer overload (src/sddmmKernel.cu:2539)
void sddmm_gpu(UIN M, UIN N, UIN K, const float* dA, const float* dB, const RPHM& rphm, float* dP, Logger& logger);
// checker: host OpenMP recomputation + checkData tolerance (src/sddmm.cu:41-59, src/host.cpp:44-76,
// include/checkData.hpp:14-30).  Verification only -- never used to produce results.
bool checkSddmm(const Matrix<float>& matrixA, const Matrix<float>& matrixB, const sparseMatrix::CSR<float>& matrixS,
                const sparseMatrix::CSR<float>& matrixP);
