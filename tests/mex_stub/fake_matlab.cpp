// fake_matlab.cpp — implementation of the fake C Matrix API in mex.h plus a small
// extern "C" driver (fm_*) that Python uses to build arguments, call mexFunction
// and read results.  TEST INFRASTRUCTURE.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mex.h"

struct mxArray_tag {
    mxClassID cls = mxDOUBLE_CLASS;
    std::vector<mwSize> dims{0, 0};
    std::vector<double> real;         // double / logical payload
    std::vector<uint64_t> u64;
    std::vector<mxArray *> cells;
    std::vector<std::string> fields;  // struct: field i of element 0 in cells[i]
    size_t numel() const {
        size_t n = 1;
        for (mwSize d : dims) n *= d;
        return n;
    }
};

static std::string g_error;
static void (*g_at_exit)(void) = nullptr;
struct MexError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

extern "C" {

double mxGetScalar(const mxArray *a) {
    if (a->cls == mxUINT64_CLASS) return a->u64.empty() ? 0.0 : (double)a->u64[0];
    return a->real.empty() ? 0.0 : a->real[0];
}
bool mxIsEmpty(const mxArray *a) { return a->numel() == 0; }
bool mxIsCell(const mxArray *a) { return a->cls == mxCELL_CLASS; }
bool mxIsStruct(const mxArray *a) { return a->cls == mxSTRUCT_CLASS; }
size_t mxGetM(const mxArray *a) { return a->dims[0]; }
size_t mxGetN(const mxArray *a) {
    size_t n = 1;
    for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i];
    return n;
}
size_t mxGetNumberOfElements(const mxArray *a) { return a->numel(); }
mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->dims.size(); }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims.data(); }
double *mxGetDoubles(const mxArray *a) { return const_cast<double *>(a->real.data()); }
void *mxGetData(const mxArray *a) {
    return a->cls == mxUINT64_CLASS ? (void *)a->u64.data() : (void *)a->real.data();
}
mxArray *mxGetCell(const mxArray *a, mwIndex i) { return i < a->cells.size() ? a->cells[i] : nullptr; }
void mxSetCell(mxArray *a, mwIndex i, mxArray *v) { a->cells.at(i) = v; }
mxArray *mxGetField(const mxArray *a, mwIndex, const char *name) {
    for (size_t i = 0; i < a->fields.size(); ++i)
        if (a->fields[i] == name) return a->cells[i];
    return nullptr;
}
void mxSetField(mxArray *a, mwIndex, const char *name, mxArray *v) {
    for (size_t i = 0; i < a->fields.size(); ++i)
        if (a->fields[i] == name) { a->cells[i] = v; return; }
}
mxArray *mxGetProperty(const mxArray *a, mwIndex i, const char *name) { return mxGetField(a, i, name); }
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity) {
    mxArray *a = new mxArray;
    a->dims = {m, n};
    a->real.assign(m * n, 0.0);
    return a;
}
mxArray *mxCreateDoubleScalar(double v) {
    mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->real[0] = v;
    return a;
}
mxArray *mxCreateLogicalScalar(bool v) {
    mxArray *a = mxCreateDoubleScalar(v ? 1.0 : 0.0);
    a->cls = mxLOGICAL_CLASS;
    return a;
}
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity) {
    mxArray *a = new mxArray;
    a->cls = cls;
    a->dims = {m, n};
    if (cls == mxUINT64_CLASS) a->u64.assign(m * n, 0);
    else a->real.assign(m * n, 0.0);
    return a;
}
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity) {
    mxArray *a = new mxArray;
    a->cls = cls;
    a->dims.assign(dims, dims + ndim);
    size_t n = 1;
    for (mwSize i = 0; i < ndim; ++i) n *= dims[i];
    if (cls == mxUINT64_CLASS) a->u64.assign(n, 0);
    else a->real.assign(n, 0.0);
    return a;
}
mxArray *mxCreateCellMatrix(mwSize m, mwSize n) {
    mxArray *a = new mxArray;
    a->cls = mxCELL_CLASS;
    a->dims = {m, n};
    a->cells.assign(m * n, nullptr);
    return a;
}
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names) {
    mxArray *a = new mxArray;
    a->cls = mxSTRUCT_CLASS;
    a->dims = {m, n};
    for (int i = 0; i < nfields; ++i) a->fields.push_back(names[i]);
    a->cells.assign(nfields, nullptr);
    return a;
}
int mexAtExit(void (*fn)(void)) {
    g_at_exit = fn;
    return 0;
}
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw MexError(std::string(id) + ": " + buf);   // MATLAB unwinds out of the MEX the same way
}

// ---- driver for Python (ctypes) ---------------------------------------------------
mxArray *fm_double(size_t ndim, const size_t *dims, const double *data) {
    mxArray *a = new mxArray;
    a->dims.assign(dims, dims + ndim);
    while (a->dims.size() < 2) a->dims.push_back(1);
    a->real.assign(data, data + a->numel());
    return a;
}
mxArray *fm_cell(size_t m, size_t n) { return mxCreateCellMatrix(m, n); }
void fm_set_cell(mxArray *c, size_t i, mxArray *v) { mxSetCell(c, i, v); }
mxArray *fm_struct(int nfields, const char **names) { return mxCreateStructMatrix(1, 1, nfields, names); }
void fm_set_field(mxArray *s, const char *name, mxArray *v) { mxSetField(s, 0, name, v); }
int fm_call(int nlhs, mxArray **plhs, int nrhs, mxArray **prhs) {
    try {
        mexFunction(nlhs, plhs, nrhs, const_cast<const mxArray **>(prhs));
        return 0;
    } catch (const MexError &e) {
        g_error = e.what();
        return 1;
    }
}
const char *fm_last_error(void) { return g_error.c_str(); }
void fm_clear_mex(void) {   // `clear mex`
    if (g_at_exit) g_at_exit();
}
size_t fm_ndim(const mxArray *a) { return a->dims.size(); }
size_t fm_dim(const mxArray *a, size_t i) { return a->dims[i]; }
int fm_class(const mxArray *a) { return (int)a->cls; }
const double *fm_doubles(const mxArray *a) { return a->real.data(); }
uint64_t fm_u64(const mxArray *a) { return a->u64.empty() ? 0 : a->u64[0]; }
mxArray *fm_get_cell(const mxArray *a, size_t i) { return mxGetCell(a, i); }
mxArray *fm_get_field(const mxArray *a, const char *name) { return mxGetField(a, 0, name); }
void fm_free(mxArray *a) {
    if (!a) return;
    for (mxArray *c : a->cells) fm_free(c);
    delete a;
}
}
