/* mex.h — FAKE MATLAB C Matrix API (test infrastructure).
 *
 * MATLAB is not available in this image.  This header plus
 * fake_matlab.cpp implement just enough of the documented C Matrix API
 * (-R2018a, interleaved complex) for p-dmpc_b200/matlab/pdmpc_b200_mex.cpp to be
 * compiled UNMODIFIED and driven from Python (tests/test_mex_shim*.py): arrays
 * are column-major double / uint64 / logical, cells, and 1x1 structs.  Written
 * from the public API documentation; contains no MATLAB code. */
#ifndef FAKE_MEX_H
#define FAKE_MEX_H
#include <stddef.h>
#include <stdint.h>

typedef size_t mwSize;
typedef size_t mwIndex;
typedef struct mxArray_tag mxArray;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxDOUBLE_CLASS = 6, mxUINT64_CLASS = 13 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;

#ifdef __cplusplus
extern "C" {
#endif
double mxGetScalar(const mxArray *a);
bool mxIsEmpty(const mxArray *a);
bool mxIsCell(const mxArray *a);
bool mxIsStruct(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
mwSize mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
double *mxGetDoubles(const mxArray *a);
void *mxGetData(const mxArray *a);
mxArray *mxGetCell(const mxArray *a, mwIndex i);
void mxSetCell(mxArray *a, mwIndex i, mxArray *v);
mxArray *mxGetField(const mxArray *a, mwIndex i, const char *name);
void mxSetField(mxArray *a, mwIndex i, const char *name, mxArray *v);
mxArray *mxGetProperty(const mxArray *a, mwIndex i, const char *name);
mxArray *mxCreateLogicalScalar(bool v);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
mxArray *mxCreateCellMatrix(mwSize m, mwSize n);
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names);
int mexAtExit(void (*fn)(void));
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
