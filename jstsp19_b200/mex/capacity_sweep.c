/* C = capacity_sweep(Y, W_zc, W_q, Mr_range, ind, scale)
 * The loop body of plot_capacity.m:45-64 (plot_ee.m:45-64) for one realization and a whole Mr range: C is numel(Mr_range) x 4 =
 * [digital BF, conventional HBF with phase shifters, conventional HBF with ZC, proposed] spectral efficiencies
 *   real(log2(det(eye(Mr) + scale * Wsel' * (Y*Y') * Wsel)))
 * Y: noiseless received block Nr x T (4th output of hbf); W_zc, W_q: createBeamformer(Nr,'ZC') / (Nr,'quantized');
 * ind = randperm(Mr_e); scale = 1/square_noise_variance*1/Nt. */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "capacity_sweep";
    gw_nargs(fn, nrhs, 6, nlhs, 1);
    const int Nr = (int)mxGetM(prhs[0]), T = (int)mxGetN(prhs[0]);
    if ((int)mxGetM(prhs[1]) != Nr || (int)mxGetN(prhs[1]) != Nr || (int)mxGetM(prhs[2]) != Nr || (int)mxGetN(prhs[2]) != Nr)
        mexErrMsgIdAndTxt("jstsp:size", "%s: W_zc and W_q must be Nr x Nr", fn);
    const int n_mr = (int)mxGetNumberOfElements(prhs[3]), n_ind = (int)mxGetNumberOfElements(prhs[4]);
    const double* mrd = gw_real(prhs[3], fn, "Mr_range");
    const double* indd = gw_real(prhs[4], fn, "ind");
    double scale = gw_scalar(prhs[5], fn, "scale");
    int* mr = (int*)mxMalloc(sizeof(int) * (n_mr ? n_mr : 1));
    int* ind = (int*)mxMalloc(sizeof(int) * (n_ind ? n_ind : 1));
    int mx = 0;
    for (int i = 0; i < n_mr; ++i) { mr[i] = (int)mrd[i]; if (mr[i] > mx) mx = mr[i]; }
    for (int i = 0; i < n_ind; ++i) ind[i] = (int)indd[i];
    if (n_ind < mx) { mxFree(mr); mxFree(ind); mexErrMsgIdAndTxt("jstsp:size", "%s: ind must hold at least max(Mr_range) entries", fn); }
    void *t0, *t1, *t2;
    const mxComplexDouble* Y = gw_complex(prhs[0], fn, "Y", &t0);
    const mxComplexDouble* Wz = gw_complex(prhs[1], fn, "W_zc", &t1);
    const mxComplexDouble* Wq = gw_complex(prhs[2], fn, "W_q", &t2);
    plhs[0] = mxCreateDoubleMatrix(n_mr, 4, mxREAL);
    double* tmp = (double*)mxMalloc(sizeof(double) * 4 * (n_mr ? n_mr : 1));
    int rc = jstsp_capacity_sweep(gw_handle(fn), JSTSP_F64, JSTSP_HOST, Nr, T, 1, n_mr, mr, Y, 0, Wz, Wq, ind, 0, &scale, tmp);
    double* out = mxGetDoubles(plhs[0]);
    for (int i = 0; i < n_mr; ++i) for (int d = 0; d < 4; ++d) out[i + (size_t)n_mr * d] = tmp[i * 4 + d];      /* column-major n_mr x 4 */
    mxFree(tmp); mxFree(mr); mxFree(ind);
    if (t0) mxFree(t0); if (t1) mxFree(t1); if (t2) mxFree(t2);
    gw_status(rc, fn);
}
