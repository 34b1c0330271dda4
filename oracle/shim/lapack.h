/* Shim "lapack.h" for the out-of-tree build of the reference (TEST INFRASTRUCTURE).
 *
 * The reference includes "lapack.h" (include/solver.h:27) and calls the LAPACK_*
 * macros (include/solver.h:100,133,249,304,360).  This image has no system LAPACK
 * headers, only an OpenBLAS 0.3.15 shared object that exports the plain Fortran
 * symbols.  This header declares exactly the five routines the reference calls.
 * It is written from the public LAPACK Fortran interface; nothing here comes
 * from the reference tree.
 */
#ifndef EMME_ORACLE_LAPACK_SHIM_H
#define EMME_ORACLE_LAPACK_SHIM_H
#include <complex>
typedef int lapack_int;
extern "C" {
void zsysv_(const char* uplo, const lapack_int* n, const lapack_int* nrhs,
            std::complex<double>* a, const lapack_int* lda, lapack_int* ipiv,
            std::complex<double>* b, const lapack_int* ldb,
            std::complex<double>* work, const lapack_int* lwork, lapack_int* info);
void zgesdd_(const char* jobz, const lapack_int* m, const lapack_int* n,
             std::complex<double>* a, const lapack_int* lda, double* s,
             std::complex<double>* u, const lapack_int* ldu,
             std::complex<double>* vt, const lapack_int* ldvt,
             std::complex<double>* work, const lapack_int* lwork, double* rwork,
             lapack_int* iwork, lapack_int* info);
void zgeqp3_(const lapack_int* m, const lapack_int* n, std::complex<double>* a,
             const lapack_int* lda, lapack_int* jpvt, std::complex<double>* tau,
             std::complex<double>* work, const lapack_int* lwork, double* rwork,
             lapack_int* info);
void ztrtrs_(const char* uplo, const char* trans, const char* diag,
             const lapack_int* n, const lapack_int* nrhs,
             const std::complex<double>* a, const lapack_int* lda,
             std::complex<double>* b, const lapack_int* ldb, lapack_int* info);
void zunmqr_(const char* side, const char* trans, const lapack_int* m,
             const lapack_int* n, const lapack_int* k,
             const std::complex<double>* a, const lapack_int* lda,
             const std::complex<double>* tau, std::complex<double>* c,
             const lapack_int* ldc, std::complex<double>* work,
             const lapack_int* lwork, lapack_int* info);
}
#define LAPACK_zsysv zsysv_
#define LAPACK_zgesdd zgesdd_
#define LAPACK_zgeqp3 zgeqp3_
#define LAPACK_ztrtrs ztrtrs_
#define LAPACK_zunmqr zunmqr_
#endif
