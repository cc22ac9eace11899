/* [H, Zbar, Ar, At, Dr, Dt] = wideband_mmwave_channel(L, Mr, Mt, Ncl, Nray, Gr, Gt)
 * drop-in for basic_system_functions/wideband_mmwave_channel.m:1.  The random draws come from MATLAB's own
 * generator through mexCallMATLAB in the reference's order (randn, randn, rand, rand per ray, .m:19-22), so a
 * seeded MATLAB session reproduces the reference's stream. */
#include "gateway_common.h"
static double draw(const char* what) {
    mxArray* o[1] = {NULL};
    if (mexCallMATLAB(1, o, 0, NULL, what) != 0 || !o[0]) mexErrMsgIdAndTxt("jstsp:rng", "wideband_mmwave_channel: %s failed", what);
    double v = mxGetScalar(o[0]);
    mxDestroyArray(o[0]);
    return v;
}
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "wideband_mmwave_channel";
    gw_nargs(fn, nrhs, 7, nlhs, 6);
    int L = (int)gw_scalar(prhs[0], fn, "L"), Mr = (int)gw_scalar(prhs[1], fn, "Mr"), Mt = (int)gw_scalar(prhs[2], fn, "Mt");
    int ncl = (int)gw_scalar(prhs[3], fn, "total_num_of_clusters"), nray = (int)gw_scalar(prhs[4], fn, "total_num_of_rays");
    int Gr = (int)gw_scalar(prhs[5], fn, "Gr"), Gt = (int)gw_scalar(prhs[6], fn, "Gt");
    int Np = ncl * nray;
    if (L < 1 || Mr < 1 || Mt < 1 || Np < 1 || Gr < 1 || Gt < 1) mexErrMsgIdAndTxt("jstsp:size", "%s: sizes must be positive", fn);
    double* normals = (double*)mxMalloc(sizeof(double) * 2 * L * Np);
    double* uniforms = (double*)mxMalloc(sizeof(double) * 2 * L * Np);
    for (int k = 0; k < L * Np; ++k) {
        normals[2 * k] = draw("randn"); normals[2 * k + 1] = draw("randn");     /* .m:19 */
        uniforms[2 * k] = draw("rand");                                          /* .m:20 */
        uniforms[2 * k + 1] = draw("rand");                                      /* .m:22 */
    }
    mwSize dH[3] = {(mwSize)Mr, (mwSize)Mt, (mwSize)L}, dAr[3] = {(mwSize)Mr, (mwSize)Np, (mwSize)L}, dAt[3] = {(mwSize)Mt, (mwSize)Np, (mwSize)L};
    mxArray* o[6];
    o[0] = mxCreateNumericArray(3, dH, mxDOUBLE_CLASS, mxCOMPLEX);
    o[1] = mxCreateDoubleMatrix(Gr, (mwSize)L * Gt, mxCOMPLEX);
    o[2] = mxCreateNumericArray(3, dAr, mxDOUBLE_CLASS, mxCOMPLEX);
    o[3] = mxCreateNumericArray(3, dAt, mxDOUBLE_CLASS, mxCOMPLEX);
    o[4] = mxCreateDoubleMatrix(Mr, Gr, mxCOMPLEX);
    o[5] = mxCreateDoubleMatrix(Mt, Gt, mxCOMPLEX);
    int rc = jstsp_wideband_mmwave_channel(gw_handle(fn), JSTSP_F64, JSTSP_HOST, L, Mr, Mt, ncl, nray, Gr, Gt, 1, normals, uniforms,
                                           mxGetComplexDoubles(o[0]), mxGetComplexDoubles(o[1]), mxGetComplexDoubles(o[2]),
                                           mxGetComplexDoubles(o[3]), mxGetComplexDoubles(o[4]), mxGetComplexDoubles(o[5]));
    mxFree(normals); mxFree(uniforms);
    int nout = nlhs > 1 ? nlhs : 1;
    for (int k = 0; k < 6; ++k) { if (k < nout) plhs[k] = o[k]; else mxDestroyArray(o[k]); }
    gw_status(rc, fn);
}
