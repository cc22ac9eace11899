/* [S, convergence_error] = sparse_admm(Htrue, OH, Dr, Dt, Imax)   drop-in for benchmark_algorithms/sparse_admm.m:1 */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "sparse_admm";
    gw_nargs(fn, nrhs, 5, nlhs, 2);
    int Mr = (int)mxGetM(prhs[1]), Mt = (int)mxGetN(prhs[1]);
    if ((int)mxGetM(prhs[2]) != Mr || (int)mxGetN(prhs[2]) != Mr || (int)mxGetM(prhs[3]) != Mt || (int)mxGetN(prhs[3]) != Mt)
        mexErrMsgIdAndTxt("jstsp:size", "%s: Dr must be Mr x Mr and Dt Mt x Mt (the reference's reshape(s,Mr,Mt) needs Gr*Gt == Mr*Mt)", fn);
    int imax = (int)gw_scalar(prhs[4], fn, "Imax");
    void *t0, *t1, *t2, *t3;
    const mxComplexDouble* Ht = gw_complex(prhs[0], fn, "Htrue", &t0);
    const mxComplexDouble* OH = gw_complex(prhs[1], fn, "OH", &t1);
    const mxComplexDouble* Dr = gw_complex(prhs[2], fn, "Dr", &t2);
    const mxComplexDouble* Dt = gw_complex(prhs[3], fn, "Dt", &t3);
    plhs[0] = mxCreateDoubleMatrix(Mr, Mt, mxCOMPLEX);
    mxArray* cv = nlhs >= 2 ? mxCreateDoubleMatrix(imax, 1, mxREAL) : NULL;
    long long ld = (long long)Mr * Mt;
    int rc = jstsp_sparse_admm(gw_handle(fn), JSTSP_F64, JSTSP_HOST, Mr, Mt, 1, imax, cv ? Ht : NULL, ld, OH, ld, Dr, 0, Dt, 0,
                               mxGetComplexDoubles(plhs[0]), ld, cv ? mxGetDoubles(cv) : NULL, imax);
    if (t0) mxFree(t0); if (t1) mxFree(t1); if (t2) mxFree(t2); if (t3) mxFree(t3);
    if (cv) plhs[1] = cv;
    gw_status(rc, fn);
}
