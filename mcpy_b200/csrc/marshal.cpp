/* Host-side helpers of the plug-in calls (CPython C API, C++17, no CUDA): list marshalling, the per-trial fast path
 * and the extxyz frame writer.
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
#include <math.h>
#include <charconv>
#include <string>

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

/* ---- extxyz frames ---------------------------------------------------------------------------------------------
 * write_xyz (mcpy/ensembles/base_ensemble.py:267-305) builds one frame per call with an f-string per atom:
 *     "{n}\n" 'energy={E:.6f}[ {extra}] Lattice="{9 x :.8f}"[ Properties=species:S:1:pos:R:3:molecule_id:I:1]\n'
 *     "{symbol} {x} {y} {z}[ {molecule_id}]\n" per atom, x / y / z as str(numpy.float64) = Python's repr(float)
 * -- 1.3 ms for 2 000 atoms, every step at the default trajectory_write_interval = 1.
 * xyz_frame() returns the same text, byte for byte: std::to_chars gives the shortest round-trip digits (what
 * repr() prints), laid out by repr()'s rule -- scientific when the decimal exponent is < -4 or >= 16, two exponent
 * digits at least, ".0" after an integral value. */
static const char* const kSymbols[] = {
    "X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar", "K", "Ca",
    "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr", "Rb", "Sr", "Y", "Zr",
    "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I", "Xe", "Cs", "Ba", "La", "Ce", "Pr", "Nd",
    "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W", "Re", "Os", "Ir", "Pt", "Au", "Hg",
    "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U", "Np", "Pu", "Am", "Cm", "Bk", "Cf", "Es", "Fm",
    "Md", "No", "Lr", "Rf", "Db", "Sg", "Bh", "Hs", "Mt", "Ds", "Rg", "Cn", "Nh", "Fl", "Mc", "Lv", "Ts", "Og"};

static void put_repr(std::string& out, double v) {
    if (isnan(v)) { out += "nan"; return; }
    if (isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
    if (v == 0.0) { out += signbit(v) ? "-0.0" : "0.0"; return; }
    char buf[40];
    const auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    const char* p = buf;
    if (*p == '-') { out += '-'; ++p; }
    // d[.ddd]e[+-]XX
    char digits[24];
    int nd = 0;
    const char* q = p;
    for (; q < res.ptr && *q != 'e'; ++q) if (*q != '.') digits[nd++] = *q;
    int ex = 0;
    {
        const char* r = q + 1;
        const bool neg = *r == '-';
        if (*r == '-' || *r == '+') ++r;
        for (; r < res.ptr; ++r) ex = 10 * ex + (*r - '0');
        if (neg) ex = -ex;
    }
    const int decpt = ex + 1;                            // value = 0.DIGITS x 10^decpt ... point after decpt digits
    if (decpt <= -4 || decpt > 16) {
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits + 1, nd - 1); }
        out += 'e';
        out += ex < 0 ? '-' : '+';
        const int ax = ex < 0 ? -ex : ex;
        char eb[8];
        const int n = snprintf(eb, sizeof eb, "%02d", ax);
        out.append(eb, n);
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out.append(digits, nd);
    } else if (decpt >= nd) {
        out.append(digits, nd);
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, decpt);
        out += '.';
        out.append(digits + decpt, nd - decpt);
    }
}

static void put_fixed(std::string& out, double v, int prec) {
    if (isnan(v)) { out += "nan"; return; }
    if (isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
    char buf[400];
    const int n = snprintf(buf, sizeof buf, "%.*f", prec, v);
    out.append(buf, n);
}

/* xyz_frame(numbers, positions, cell, energy, extra, molecule_id | None) -> str, or None when the arrays are not
 * (n,) integer / (n, 3) float64 / 9 float64 / (n,) integer in plain layout (the caller then formats in Python). */
static PyObject* xyz_frame(PyObject* self, PyObject* args) {
    PyObject *oz, *op, *oc, *om;
    double energy;
    const char* extra;
    Py_ssize_t extra_len;
    if (!PyArg_ParseTuple(args, "OOOds#O", &oz, &op, &oc, &energy, &extra, &extra_len, &om)) return NULL;
    if (!plain(op, NPY_FLOAT64) || PyArray_NDIM((PyArrayObject*)op) != 2 || PyArray_DIM((PyArrayObject*)op, 1) != 3) Py_RETURN_NONE;
    const npy_intp n = PyArray_DIM((PyArrayObject*)op, 0);
    auto int_width = [](PyObject* o) -> int {
        if (!o || !PyArray_Check(o)) return 0;
        const int t = PyArray_TYPE((PyArrayObject*)o);
        const int w = (t == NPY_INT64 || t == NPY_LONGLONG || (t == NPY_LONG && sizeof(long) == 8)) ? 8
                    : (t == NPY_INT32 || t == NPY_INT || (t == NPY_LONG && sizeof(long) == 4)) ? 4 : 0;
        return (w && plain(o, t)) ? w : 0;
    };
    const int zw = int_width(oz);
    if (!zw || PyArray_NDIM((PyArrayObject*)oz) != 1 || PyArray_DIM((PyArrayObject*)oz, 0) != n) Py_RETURN_NONE;
    if (!plain(oc, NPY_FLOAT64) || PyArray_SIZE((PyArrayObject*)oc) != 9) Py_RETURN_NONE;
    int mw = 0;
    if (om != Py_None) {
        mw = int_width(om);
        if (!mw || PyArray_NDIM((PyArrayObject*)om) != 1 || PyArray_DIM((PyArrayObject*)om, 0) != n) Py_RETURN_NONE;
    }
    const double* pos = (const double*)PyArray_DATA((PyArrayObject*)op);
    const double* cell = (const double*)PyArray_DATA((PyArrayObject*)oc);
    const void* zs = PyArray_DATA((PyArrayObject*)oz);
    const void* ms = mw ? PyArray_DATA((PyArrayObject*)om) : NULL;
    auto at = [](const void* base, int w, npy_intp i) -> long long {
        return w == 8 ? (long long)((const int64_t*)base)[i] : (long long)((const int32_t*)base)[i];
    };
    for (npy_intp i = 0; i < n; ++i) {
        const long long z = at(zs, zw, i);
        if (z < 0 || z >= (long long)(sizeof kSymbols / sizeof kSymbols[0])) Py_RETURN_NONE;
    }
    std::string out;
    out.reserve(64 + 160 + (size_t)n * 72);
    out += std::to_string((long long)n);
    out += "\nenergy=";
    put_fixed(out, energy, 6);
    if (extra_len) { out += ' '; out.append(extra, (size_t)extra_len); }
    out += " Lattice=\"";
    for (int k = 0; k < 9; ++k) { if (k) out += ' '; put_fixed(out, cell[k], 8); }
    out += '"';
    if (mw) out += " Properties=species:S:1:pos:R:3:molecule_id:I:1";
    out += '\n';
    for (npy_intp i = 0; i < n; ++i) {
        out += kSymbols[at(zs, zw, i)];
        out += ' '; put_repr(out, pos[3 * i]);
        out += ' '; put_repr(out, pos[3 * i + 1]);
        out += ' '; put_repr(out, pos[3 * i + 2]);
        if (mw) { out += ' '; out += std::to_string(at(ms, mw, i)); }
        out += '\n';
    }
    return PyUnicode_FromStringAndSize(out.data(), (Py_ssize_t)out.size());
}

static PyMethodDef methods[] = {
    {"xyz_frame", xyz_frame, METH_VARARGS, "one extxyz frame as text, byte for byte what the reference's write_xyz builds"},
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
