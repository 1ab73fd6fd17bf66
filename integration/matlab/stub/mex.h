/* Minimal stand-in for MATLAB's mex.h: declarations only, enough to compile-check
 * integration/matlab/nsagp_mex.cpp where MATLAB is not installed.  Never linked. */
#ifndef NSAGP_STUB_MEX_H
#define NSAGP_STUB_MEX_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
#ifdef __cplusplus
extern "C" {
#endif
bool mxIsStruct(const mxArray*);
bool mxIsDouble(const mxArray*);
bool mxIsComplex(const mxArray*);
bool mxIsEmpty(const mxArray*);
bool mxIsChar(const mxArray*);
mxArray* mxGetField(const mxArray*, mwSize, const char*);
double* mxGetDoubles(const mxArray*);
double mxGetScalar(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
mxArray* mxCreateString(const char*);
int mxAddField(mxArray*, const char*);
void mxSetField(mxArray*, mwSize, const char*, mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
void mxDestroyArray(mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexFunction(int, mxArray*[], int, const mxArray*[]);
#ifdef __cplusplus
}
#endif
#endif
