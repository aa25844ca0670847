/* Host-side marshalling for the batched plug-in call (CPython C API, no CUDA).
 *
 * B200Calculator.get_potential_energies(atoms_list) (the reference's BaseCalculator.get_potential_energies,
 * base_calculator.py:24-27, as BatchedReplicaExchange._batched_single_move calls it once per step,
 * batched_replica_exchange.py:277-278) hands the engine one pointer per configuration: the positions and numbers
 * arrays are read where ase.Atoms keeps them.  Collecting 64 x 4 attributes and addresses from Python costs
 * ~135 us per call -- more than the compare + launch + readback it feeds (~110 us) -- so the walk over the list is
 * done here: rows() fills the pointer / size tables and copies cell and pbc of every configuration, ~10 us for 64.
 *
 * rows(atoms_list, pos_ptrs, num_ptrs, n_atoms, cells, pbcs) -> 0 (numbers are int32), 1 (int64), or -1 when
 * some configuration cannot be read in place (the caller then takes the generic numpy path).
 *   pos_ptrs, num_ptrs : writable uintp[R];  n_atoms : writable int64[R];
 *   cells : writable float64[R, 9];  pbcs : writable uint8[R, 3]
 * A configuration is anything with .positions (C-contiguous float64 (n, 3)), .numbers (C-contiguous int32 or
 * int64 (n,)), .cell ((3, 3) float64, or an object whose .array is one: ase.cell.Cell) and .pbc (3 bools);
 * ase.Atoms is read through its .arrays dict.  The arrays stay owned by the configurations: the caller must keep
 * atoms_list alive (and unmodified) until the engine call that consumes the tables returns.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <stdint.h>
#include <string.h>

static PyObject *s_arrays, *s_positions, *s_numbers, *s_cell, *s_pbc, *s_array, *s_cellobj, *s_pbc_priv;

/* a C-contiguous, aligned ndarray of the given type number? */
static int plain(PyObject* o, int typenum) {
    if (!o || !PyArray_Check(o)) return 0;
    PyArrayObject* a = (PyArrayObject*)o;
    return PyArray_TYPE(a) == typenum && PyArray_IS_C_CONTIGUOUS(a) && PyArray_ISALIGNED(a);
}

/* attribute lookup that leaves no error behind (and does not build an AttributeError for a missing one: ~1 us) */
static PyObject* attr(PyObject* o, PyObject* name) {
    PyObject* v = NULL;
#if PY_VERSION_HEX >= 0x030D0000
    if (PyObject_GetOptionalAttr(o, name, &v) < 0) PyErr_Clear();
#else
    if (_PyObject_LookupAttr(o, name, &v) < 0) PyErr_Clear();
#endif
    return v;
}

static PyObject* rows(PyObject* self, PyObject* args) {
    PyObject *lst, *o_pp, *o_zp, *o_n, *o_cells, *o_pbcs;
    if (!PyArg_ParseTuple(args, "OOOOOO", &lst, &o_pp, &o_zp, &o_n, &o_cells, &o_pbcs)) return NULL;
    PyObject* seq = PySequence_Fast(lst, "atoms_list must be a sequence");
    if (!seq) return NULL;
    const Py_ssize_t R = PySequence_Fast_GET_SIZE(seq);
    if (!plain(o_pp, NPY_UINTP) || !plain(o_zp, NPY_UINTP) || !plain(o_n, NPY_INT64) || !plain(o_cells, NPY_FLOAT64) ||
        !plain(o_pbcs, NPY_UINT8) || PyArray_SIZE((PyArrayObject*)o_pp) < R || PyArray_SIZE((PyArrayObject*)o_zp) < R ||
        PyArray_SIZE((PyArrayObject*)o_n) < R || PyArray_SIZE((PyArrayObject*)o_cells) < 9 * R ||
        PyArray_SIZE((PyArrayObject*)o_pbcs) < 3 * R) {
        Py_DECREF(seq);
        PyErr_SetString(PyExc_ValueError, "marshal.rows: output tables must be contiguous uintp[R], uintp[R], int64[R], "
                                          "float64[R, 9], uint8[R, 3]");
        return NULL;
    }
    uintptr_t* pp = (uintptr_t*)PyArray_DATA((PyArrayObject*)o_pp);
    uintptr_t* zp = (uintptr_t*)PyArray_DATA((PyArrayObject*)o_zp);
    int64_t* n_of = (int64_t*)PyArray_DATA((PyArrayObject*)o_n);
    double* cells = (double*)PyArray_DATA((PyArrayObject*)o_cells);
    uint8_t* pbcs = (uint8_t*)PyArray_DATA((PyArrayObject*)o_pbcs);
    int ztype = -1, bad = 0;
    for (Py_ssize_t r = 0; r < R && !bad; ++r) {
        PyObject* a = PySequence_Fast_GET_ITEM(seq, r);          /* borrowed */
        PyObject *p = NULL, *z = NULL, *c = NULL, *b = NULL;
        /* ase.Atoms keeps positions / numbers in its .arrays dict, the cell matrix in ._cellobj.array and the
           periodicity flags in ._pbc; the public properties return exactly these objects */
        PyObject* arrays = attr(a, s_arrays);
        if (arrays && PyDict_CheckExact(arrays)) {
            p = PyDict_GetItemWithError(arrays, s_positions);
            z = PyDict_GetItemWithError(arrays, s_numbers);
            Py_XINCREF(p); Py_XINCREF(z);
            PyErr_Clear();
            PyObject* co = attr(a, s_cellobj);
            if (co) { c = attr(co, s_array); Py_DECREF(co); }
            b = attr(a, s_pbc_priv);
        }
        Py_XDECREF(arrays);
        if (!p) p = attr(a, s_positions);
        if (!z) z = attr(a, s_numbers);
        if (!c) {
            c = attr(a, s_cell);
            if (c && !PyArray_Check(c)) {                        /* ase.cell.Cell: the matrix is .array */
                PyObject* inner = attr(c, s_array);
                Py_DECREF(c);
                c = inner;
            }
        }
        if (!b) b = attr(a, s_pbc);
        if (!plain(p, NPY_FLOAT64) || PyArray_NDIM((PyArrayObject*)p) != 2 || PyArray_DIM((PyArrayObject*)p, 1) != 3) bad = 1;
        if (!bad) {
            const npy_intp n = PyArray_DIM((PyArrayObject*)p, 0);
            n_of[r] = (int64_t)n;
            pp[r] = n ? (uintptr_t)PyArray_DATA((PyArrayObject*)p) : 0;
            const int zt = z && PyArray_Check(z) ? PyArray_TYPE((PyArrayObject*)z) : -1;
            const int w = (zt == NPY_INT64 || (zt == NPY_LONG && sizeof(long) == 8) || zt == NPY_LONGLONG) ? 8
                        : (zt == NPY_INT32 || (zt == NPY_LONG && sizeof(long) == 4) || zt == NPY_INT) ? 4 : 0;
            if (!w || !plain(z, zt) || PyArray_NDIM((PyArrayObject*)z) != 1 || PyArray_DIM((PyArrayObject*)z, 0) != n ||
                (ztype >= 0 && w != ztype)) bad = 1;
            else { ztype = w; zp[r] = n ? (uintptr_t)PyArray_DATA((PyArrayObject*)z) : 0; }
        }
        if (!bad) {
            if (!plain(c, NPY_FLOAT64) || PyArray_SIZE((PyArrayObject*)c) != 9) bad = 1;
            else memcpy(cells + 9 * r, PyArray_DATA((PyArrayObject*)c), 72);
        }
        if (!bad) {
            if (!b || !PyArray_Check(b) || PyArray_ITEMSIZE((PyArrayObject*)b) != 1 || PyArray_SIZE((PyArrayObject*)b) != 3 ||
                !PyArray_IS_C_CONTIGUOUS((PyArrayObject*)b)) bad = 1;
            else {
                const uint8_t* s = (const uint8_t*)PyArray_DATA((PyArrayObject*)b);
                pbcs[3 * r] = s[0] != 0; pbcs[3 * r + 1] = s[1] != 0; pbcs[3 * r + 2] = s[2] != 0;
            }
        }
        Py_XDECREF(p); Py_XDECREF(z); Py_XDECREF(c); Py_XDECREF(b);
    }
    Py_DECREF(seq);
    if (PyErr_Occurred()) PyErr_Clear();
    return PyLong_FromLong(bad ? -1 : (ztype == 8 ? 1 : 0));
}

/* One configuration: positions / numbers / cell / pbc of `a` read in place (0), or -1. */
typedef struct { const double* pos; const void* num; int64_t n; int zwidth; double cell[9]; uint8_t pbc[3]; } OneConfig;

static int read_one(PyObject* a, OneConfig* o) {
    PyObject *p = NULL, *z = NULL, *c = NULL, *b = NULL;
    int bad = 0;
    PyObject* arrays = attr(a, s_arrays);
    if (arrays && PyDict_CheckExact(arrays)) {
        p = PyDict_GetItemWithError(arrays, s_positions);
        z = PyDict_GetItemWithError(arrays, s_numbers);
        Py_XINCREF(p); Py_XINCREF(z);
        PyErr_Clear();
        PyObject* co = attr(a, s_cellobj);
        if (co) { c = attr(co, s_array); Py_DECREF(co); }
        b = attr(a, s_pbc_priv);
    }
    Py_XDECREF(arrays);
    if (!p) p = attr(a, s_positions);
    if (!z) z = attr(a, s_numbers);
    if (!c) {
        c = attr(a, s_cell);
        if (c && !PyArray_Check(c)) { PyObject* inner = attr(c, s_array); Py_DECREF(c); c = inner; }
    }
    if (!b) b = attr(a, s_pbc);
    if (!plain(p, NPY_FLOAT64) || PyArray_NDIM((PyArrayObject*)p) != 2 || PyArray_DIM((PyArrayObject*)p, 1) != 3) bad = 1;
    if (!bad) {
        const npy_intp n = PyArray_DIM((PyArrayObject*)p, 0);
        o->n = (int64_t)n;
        o->pos = n ? (const double*)PyArray_DATA((PyArrayObject*)p) : NULL;
        const int zt = z && PyArray_Check(z) ? PyArray_TYPE((PyArrayObject*)z) : -1;
        const int w = (zt == NPY_INT64 || (zt == NPY_LONG && sizeof(long) == 8) || zt == NPY_LONGLONG) ? 8
                    : (zt == NPY_INT32 || (zt == NPY_LONG && sizeof(long) == 4) || zt == NPY_INT) ? 4 : 0;
        if (!w || !plain(z, zt) || PyArray_NDIM((PyArrayObject*)z) != 1 || PyArray_DIM((PyArrayObject*)z, 0) != n) bad = 1;
        else { o->zwidth = w; o->num = n ? PyArray_DATA((PyArrayObject*)z) : NULL; }
    }
    if (!bad) {
        if (!plain(c, NPY_FLOAT64) || PyArray_SIZE((PyArrayObject*)c) != 9) bad = 1;
        else memcpy(o->cell, PyArray_DATA((PyArrayObject*)c), 72);
    }
    if (!bad) {
        if (!b || !PyArray_Check(b) || PyArray_ITEMSIZE((PyArrayObject*)b) != 1 || PyArray_SIZE((PyArrayObject*)b) != 3 ||
            !PyArray_IS_C_CONTIGUOUS((PyArrayObject*)b)) bad = 1;
        else {
            const uint8_t* s = (const uint8_t*)PyArray_DATA((PyArrayObject*)b);
            o->pbc[0] = s[0] != 0; o->pbc[1] = s[1] != 0; o->pbc[2] = s[2] != 0;
        }
    }
    /* the arrays stay alive through `a` (the caller holds it, and the GIL, until the engine call returns) */
    Py_XDECREF(p); Py_XDECREF(z); Py_XDECREF(c); Py_XDECREF(b);
    if (PyErr_Occurred()) PyErr_Clear();
    return bad ? -1 : 0;
}

typedef int (*tracked_fn)(void*, const double*, const void*, int, const double*, const uint8_t*, double*, int*);

/* tracked_energy(fn_i32, fn_i64, engine, atoms) -> (status, energy, path, n_atoms) or None (cannot read `atoms` in place).
 * fn_i32 / fn_i64: addresses of mcpyb_tracked_energy_host / mcpyb_tracked_energy_z64_host (include/mcpy_b200.h),
 * engine: the mcpyb_engine handle.  The drop-in call of a Monte-Carlo chain -- one per trial -- without the ~5 us
 * of numpy / ctypes marshalling around a ~9 us engine call. */
static PyObject* tracked_energy(PyObject* self, PyObject* args) {
    PyObject *o32, *o64, *oh, *atoms;
    if (!PyArg_ParseTuple(args, "OOOO", &o32, &o64, &oh, &atoms)) return NULL;
    void* f32 = PyLong_AsVoidPtr(o32);
    void* f64 = PyLong_AsVoidPtr(o64);
    void* h = PyLong_AsVoidPtr(oh);
    if (PyErr_Occurred()) return NULL;
    if (!f32 || !f64 || !h) { PyErr_SetString(PyExc_ValueError, "marshal.tracked_energy: null function or engine"); return NULL; }
    OneConfig c;
    if (read_one(atoms, &c) != 0 || c.n > 0x7fffffff) Py_RETURN_NONE;
    double e = 0.0;
    int path = 0;
    const int rc = ((tracked_fn)(c.zwidth == 8 ? f64 : f32))(h, c.pos, c.num, (int)c.n, c.cell, c.pbc, &e, &path);
    return Py_BuildValue("(idiL)", rc, e, path, (long long)c.n);
}

static PyMethodDef methods[] = {
    {"tracked_energy", tracked_energy, METH_VARARGS, "one tracked energy call on an Atoms-like object"},
    {"rows", rows, METH_VARARGS, "fill the per-configuration pointer tables of a batched engine call"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_marshal", "host-side marshalling for mcpy_b200", -1, methods};

PyMODINIT_FUNC PyInit__marshal(void) {
    s_arrays = PyUnicode_InternFromString("arrays");
    s_positions = PyUnicode_InternFromString("positions");
    s_numbers = PyUnicode_InternFromString("numbers");
    s_cell = PyUnicode_InternFromString("cell");
    s_pbc = PyUnicode_InternFromString("pbc");
    s_array = PyUnicode_InternFromString("array");
    s_cellobj = PyUnicode_InternFromString("_cellobj");
    s_pbc_priv = PyUnicode_InternFromString("_pbc");
    import_array();
    return PyModule_Create(&module);
}
