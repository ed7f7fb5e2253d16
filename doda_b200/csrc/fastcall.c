/* fastcall.c — CPython binding of the per-layer hot entry points of libb200sparse.so.
 *
 * The whole C ABI is bound through ctypes (doda_b200/_lib.py, prototypes parsed from include/b200sparse.h).  ctypes
 * costs 4-8 us per call for 16-argument functions, and one training step of DODA's U-Net makes ~550 such calls
 * (71 convs x fwd/dgrad/wgrad, 65 BN x fwd/bwd) against ~14 ms of GPU work -- the step was host-bound on it.  This
 * module calls the SAME exported functions with METH_FASTCALL argument passing (~0.5 us per call): pointers and
 * sizes as Python ints (None = NULL), nothing else.  It adds no functionality and keeps no state. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include "../../include/b200sparse.h"

static inline void* P(PyObject* o) { return o == Py_None ? NULL : PyLong_AsVoidPtr(o); }
static inline long long I(PyObject* o) { return PyLong_AsLongLong(o); }
static inline double F(PyObject* o) { return PyFloat_AsDouble(o); }

#define NARGS(n)                                                                                        \
    if (nargs != (n)) {                                                                                 \
        PyErr_Format(PyExc_TypeError, "%s expects %d arguments, got %zd", __func__, (n), nargs);        \
        return NULL;                                                                                    \
    }
#define CHECKED(call)                     \
    if (PyErr_Occurred()) return NULL;    \
    return PyLong_FromLong((long)(call));

/* b200sp_gather_gemm(in, n_in, Cin, W, wflags, tab, orow, rowmask, K, out, n_out, Cout, accumulate, ws, ws_bytes, stream) */
static PyObject* f_gather_gemm(PyObject* self, PyObject* const* a, Py_ssize_t nargs) {
    NARGS(16)
    const float* in = P(a[0]); long long n_in = I(a[1]); int Cin = (int)I(a[2]); const float* W = P(a[3]);
    int wflags = (int)I(a[4]); const int32_t* tab = P(a[5]); const int32_t* orow = P(a[6]); const int32_t* rowmask = P(a[7]);
    int K = (int)I(a[8]); float* out = P(a[9]); long long n_out = I(a[10]); int Cout = (int)I(a[11]);
    int accumulate = (int)I(a[12]); void* ws = P(a[13]); long long ws_bytes = I(a[14]); void* stream = P(a[15]);
    CHECKED(b200sp_gather_gemm(in, n_in, Cin, W, wflags, tab, orow, rowmask, K, out, n_out, Cout, accumulate, ws, ws_bytes, stream))
}

/* b200sp_gather_gemm_pairs(in, Cin, W, wflags, pin, pout, pairnum, n_upper, K, pstride, out, Cout, accumulate, ws, ws_bytes, stream) */
static PyObject* f_gather_gemm_pairs(PyObject* self, PyObject* const* a, Py_ssize_t nargs) {
    NARGS(16)
    const float* in = P(a[0]); int Cin = (int)I(a[1]); const float* W = P(a[2]); int wflags = (int)I(a[3]);
    const int32_t* pin = P(a[4]); const int32_t* pout = P(a[5]); const int32_t* pairnum = P(a[6]); long long n_upper = I(a[7]);
    int K = (int)I(a[8]); long long pstride = I(a[9]); float* out = P(a[10]); int Cout = (int)I(a[11]);
    int accumulate = (int)I(a[12]); void* ws = P(a[13]); long long ws_bytes = I(a[14]); void* stream = P(a[15]);
    CHECKED(b200sp_gather_gemm_pairs(in, Cin, W, wflags, pin, pout, pairnum, n_upper, K, pstride, out, Cout, accumulate, ws, ws_bytes, stream))
}

/* b200sp_wgrad(a, Ca, b, Cb, pa, pb, pairnum, n_upper, K, pstride, dW, stream) */
static PyObject* f_wgrad(PyObject* self, PyObject* const* a, Py_ssize_t nargs) {
    NARGS(12)
    const float* x = P(a[0]); int Ca = (int)I(a[1]); const float* g = P(a[2]); int Cb = (int)I(a[3]);
    const int32_t* pa = P(a[4]); const int32_t* pb = P(a[5]); const int32_t* pairnum = P(a[6]); long long n_upper = I(a[7]);
    int K = (int)I(a[8]); long long pstride = I(a[9]); float* dW = P(a[10]); void* stream = P(a[11]);
    CHECKED(b200sp_wgrad(x, Ca, g, Cb, pa, pb, pairnum, n_upper, K, pstride, dW, stream))
}

/* b200sp_wgrad_table(a, Ca, g, Cb, tab, orow, rowmask, n_rows, K, dW, stream) */
static PyObject* f_wgrad_table(PyObject* self, PyObject* const* a, Py_ssize_t nargs) {
    NARGS(11)
    const float* x = P(a[0]); int Ca = (int)I(a[1]); const float* g = P(a[2]); int Cb = (int)I(a[3]);
    const int32_t* tab = P(a[4]); const int32_t* orow = P(a[5]); const int32_t* rowmask = P(a[6]); long long n_rows = I(a[7]);
    int K = (int)I(a[8]); float* dW = P(a[9]); void* stream = P(a[10]);
    CHECKED(b200sp_wgrad_table(x, Ca, g, Cb, tab, orow, rowmask, n_rows, K, dW, stream))
}

/* b200sp_bn_fwd_train(x, M, C, w, b, eps, relu, y, mean, invstd, running_mean, running_var, momentum, nbt, ws, ws_bytes, stream) */
static PyObject* f_bn_fwd_train(PyObject* self, PyObject* const* a, Py_ssize_t nargs) {
    NARGS(17)
    const float* x = P(a[0]); long long M = I(a[1]); int C = (int)I(a[2]); const float* w = P(a[3]); const float* b = P(a[4]);
    float eps = (float)F(a[5]); int relu = (int)I(a[6]); float* y = P(a[7]); float* mean = P(a[8]); float* invstd = P(a[9]);
    float* rm = P(a[10]); float* rv = P(a[11]); float momentum = (float)F(a[12]); void* nbt = P(a[13]);
    void* ws = P(a[14]); long long ws_bytes = I(a[15]); void* stream = P(a[16]);
    CHECKED(b200sp_bn_fwd_train(x, M, C, w, b, eps, relu, y, mean, invstd, rm, rv, momentum, nbt, ws, ws_bytes, stream))
}

/* b200sp_bn_bwd(x, dy, M, C, w, b, mean, invstd, relu, dx, dw, db, ws, ws_bytes, stream) */
static PyObject* f_bn_bwd(PyObject* self, PyObject* const* a, Py_ssize_t nargs) {
    NARGS(15)
    const float* x = P(a[0]); const float* dy = P(a[1]); long long M = I(a[2]); int C = (int)I(a[3]); const float* w = P(a[4]);
    const float* b = P(a[5]); const float* mean = P(a[6]); const float* invstd = P(a[7]); int relu = (int)I(a[8]);
    float* dx = P(a[9]); float* dw = P(a[10]); float* db = P(a[11]); void* ws = P(a[12]); long long ws_bytes = I(a[13]);
    void* stream = P(a[14]);
    CHECKED(b200sp_bn_bwd(x, dy, M, C, w, b, mean, invstd, relu, dx, dw, db, ws, ws_bytes, stream))
}

/* b200sp_stream_fork / b200sp_stream_join (main_stream, side_stream, event) */
static PyObject* f_fork(PyObject* self, PyObject* const* a, Py_ssize_t nargs) {
    NARGS(3)
    void* m = P(a[0]); void* s = P(a[1]); void* e = P(a[2]);
    CHECKED(b200sp_stream_fork(m, s, e))
}
static PyObject* f_join(PyObject* self, PyObject* const* a, Py_ssize_t nargs) {
    NARGS(3)
    void* m = P(a[0]); void* s = P(a[1]); void* e = P(a[2]);
    CHECKED(b200sp_stream_join(m, s, e))
}

/* b200sp_conv_layer_fwd / _bwd: every argument becomes one int64 (None -> 0, float -> its double bit pattern) */
#include <string.h>
#define LAYER_MAXARGS 48
static PyObject* layer_call(PyObject* const* a, Py_ssize_t nargs, int (*fn)(const int64_t*, int)) {
    int64_t v[LAYER_MAXARGS];
    if (nargs > LAYER_MAXARGS) {
        PyErr_SetString(PyExc_TypeError, "too many arguments for a layer call");
        return NULL;
    }
    for (Py_ssize_t i = 0; i < nargs; ++i) {
        PyObject* o = a[i];
        if (o == Py_None) {
            v[i] = 0;
        } else if (PyFloat_CheckExact(o)) {
            double d = PyFloat_AS_DOUBLE(o);
            memcpy(&v[i], &d, sizeof(d));
        } else {
            v[i] = (int64_t)PyLong_AsUnsignedLongLongMask(o);
        }
    }
    if (PyErr_Occurred()) return NULL;
    return PyLong_FromLong((long)fn(v, (int)nargs));
}
static PyObject* f_layer_fwd(PyObject* self, PyObject* const* a, Py_ssize_t nargs) { return layer_call(a, nargs, b200sp_conv_layer_fwd); }
static PyObject* f_layer_bwd(PyObject* self, PyObject* const* a, Py_ssize_t nargs) { return layer_call(a, nargs, b200sp_conv_layer_bwd); }

static PyMethodDef methods[] = {
    {"layer_fwd", (PyCFunction)(void (*)(void))f_layer_fwd, METH_FASTCALL, "b200sp_conv_layer_fwd"},
    {"layer_bwd", (PyCFunction)(void (*)(void))f_layer_bwd, METH_FASTCALL, "b200sp_conv_layer_bwd"},
    {"fork", (PyCFunction)(void (*)(void))f_fork, METH_FASTCALL, "b200sp_stream_fork"},
    {"join", (PyCFunction)(void (*)(void))f_join, METH_FASTCALL, "b200sp_stream_join"},
    {"gather_gemm", (PyCFunction)(void (*)(void))f_gather_gemm, METH_FASTCALL, "b200sp_gather_gemm"},
    {"gather_gemm_pairs", (PyCFunction)(void (*)(void))f_gather_gemm_pairs, METH_FASTCALL, "b200sp_gather_gemm_pairs"},
    {"wgrad", (PyCFunction)(void (*)(void))f_wgrad, METH_FASTCALL, "b200sp_wgrad"},
    {"wgrad_table", (PyCFunction)(void (*)(void))f_wgrad_table, METH_FASTCALL, "b200sp_wgrad_table"},
    {"bn_fwd_train", (PyCFunction)(void (*)(void))f_bn_fwd_train, METH_FASTCALL, "b200sp_bn_fwd_train"},
    {"bn_bwd", (PyCFunction)(void (*)(void))f_bn_bwd, METH_FASTCALL, "b200sp_bn_bwd"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_b200fast", "fast-call binding of libb200sparse hot entry points", -1, methods};

PyMODINIT_FUNC PyInit__b200fast(void) { return PyModule_Create(&moddef); }
