/*
 * mgic_chf.h -- link-time drop-ins for the reference's Chombo-Fortran kernels.
 *
 * libmgic_b200.so exports these symbols with exactly the names and argument
 * lists the ChF preprocessor generates (FORTRAN_NAME(UPPER, lower) == lower-case
 * + trailing underscore; everything by pointer), so that the reference's
 * VariableCoeffPoissonOperator.cpp / SetLevelData.cpp link against this library
 * instead of the objects compiled from the .ChF files:
 *
 *   gsrbhelmholtzvc3d_  replaces Source/VariableCoeffPoissonOperatorF.ChF:56-139
 *                       (prototype Source/VariableCoeffPoissonOperatorF_F.H:107-117)
 *   vccomputeop3d_      replaces ...F.ChF:181-237   (prototype ..._F.H:233-241)
 *   vccomputeres3d_     replaces ...F.ChF:283-339   (prototype ..._F.H:359-368)
 *   restrictresvc3d_    replaces ...F.ChF:379-437   (prototype ..._F.H:488-497)
 *   getlaplacianpsif_   replaces Source/SetLevelDataF.ChF:15-58  (prototype Source/SetLevelDataF_F.H:15-19)
 *   getrhogradphif_     replaces Source/SetLevelDataF.ChF:65-103 (prototype Source/SetLevelDataF_F.H:43-47)
 *   prolong_            replaces [Chombo 3.2] AMRPoissonOpF.ChF PROLONG (used by AMRPoissonOp::prolongIncrement)
 *
 * Semantics: host pointers, caller owns all memory, callee keeps nothing; each
 * call stages the FABs to HBM, runs the sm_100a kernel and copies the result
 * back (this boundary exists for parity testing and link-time drop-in; the
 * timed path is the device-resident mgic_* API of mgic.h).  On error the
 * routines call abort() like the Fortran MAYDAYERROR().  A CUDA device is
 * required -- there is no CPU fallback.
 *
 * Argument macros as expanded by [Chombo] FORT_PROTO.H in 3D:
 *   CHFp_FRA(a)   -> Real* a, const int* ialo0, ialo1, ialo2, iahi0, iahi1, iahi2, const int* nacomp
 *   CHFp_FRA1(a)  -> same without nacomp
 *   CHFp_BOX(b)   -> const int* iblo0, iblo1, iblo2, ibhi0, ibhi1, ibhi2
 */
#ifndef MGIC_CHF_H
#define MGIC_CHF_H

#ifdef __cplusplus
extern "C" {
#endif

#define MGIC_FRA(a)  double *a, const int *i##a##lo0, const int *i##a##lo1, const int *i##a##lo2, \
                     const int *i##a##hi0, const int *i##a##hi1, const int *i##a##hi2, const int *n##a##comp
#define MGIC_CFRA(a) const double *a, const int *i##a##lo0, const int *i##a##lo1, const int *i##a##lo2, \
                     const int *i##a##hi0, const int *i##a##hi1, const int *i##a##hi2, const int *n##a##comp
#define MGIC_FRA1(a)  double *a, const int *i##a##lo0, const int *i##a##lo1, const int *i##a##lo2, \
                      const int *i##a##hi0, const int *i##a##hi1, const int *i##a##hi2
#define MGIC_CFRA1(a) const double *a, const int *i##a##lo0, const int *i##a##lo1, const int *i##a##lo2, \
                      const int *i##a##hi0, const int *i##a##hi1, const int *i##a##hi2
#define MGIC_BOX(b)  const int *i##b##lo0, const int *i##b##lo1, const int *i##b##lo2, \
                     const int *i##b##hi0, const int *i##b##hi1, const int *i##b##hi2

void gsrbhelmholtzvc3d_(MGIC_FRA(dpsi), MGIC_CFRA(rhs), MGIC_BOX(region), const double *dx, const double *alpha,
                        MGIC_CFRA(aCoef), const double *beta, MGIC_CFRA(bCoef), MGIC_CFRA(lambda),
                        const int *redBlack);
void vccomputeop3d_(MGIC_FRA(lofdpsi), MGIC_CFRA(dpsi), const double *alpha, MGIC_CFRA(aCoef), const double *beta,
                    MGIC_CFRA(bCoef), MGIC_BOX(region), const double *dx);
void vccomputeres3d_(MGIC_FRA(res), MGIC_CFRA(dpsi), MGIC_CFRA(rhs), const double *alpha, MGIC_CFRA(aCoef),
                     const double *beta, MGIC_CFRA(bCoef), MGIC_BOX(region), const double *dx);
void restrictresvc3d_(MGIC_FRA(res), MGIC_CFRA(dpsi), MGIC_CFRA(rhs), const double *alpha, MGIC_CFRA(aCoef),
                      const double *beta, MGIC_CFRA(bCoef), MGIC_BOX(region), const double *dx);
void getlaplacianpsif_(MGIC_FRA1(l_of_psi), MGIC_CFRA1(psi), const double *dx, MGIC_BOX(box));
void getrhogradphif_(MGIC_FRA1(rho_grad_phi), MGIC_CFRA1(phi), const double *dx, MGIC_BOX(box));
void prolong_(MGIC_FRA(phi), MGIC_CFRA(coarse), MGIC_BOX(region), const int *m);

#ifdef __cplusplus
}
#endif
#endif /* MGIC_CHF_H */
