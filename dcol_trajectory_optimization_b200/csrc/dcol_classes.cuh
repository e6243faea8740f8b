/*
 * dcol_classes.cuh — the compile-time specialisations of the pair solver and how a shape maps to one.
 *
 * The reference dispatches on isinstance for every pair it evaluates
 * (primitives/problem_matrices.py:255-364).  Here a shape is classified ONCE, when the shape table
 * is created: its class fixes the primitive kind and, for the face counts that occur in the
 * reference's scenes (axis-aligned boxes from create_rect_prism, other 6-face polytopes, the 8-face polytope of
 * systems/polytopes.jld2, the 5-gon of create_n_sided(5, .)), the number of half-spaces, so that a
 * pair of classes selects one fully unrolled, register-resident kernel.  Other face counts
 * (1..DCOL_MAX_FACES) fall to the runtime-count specialisations of the same code.
 */
#ifndef DCOL_CLASSES_CUH_
#define DCOL_CLASSES_CUH_

#include "dcol_solver.cuh"

namespace dcol {

enum {
    CLS_POLY6 = 0, CLS_POLY8, CLS_POLYN, CLS_CAPSULE, CLS_CYLINDER, CLS_CONE, CLS_SPHERE, CLS_PGON5, CLS_PGONN, CLS_BOX,
    CLS_ELLIPSOID, N_CLS
};

template <int CLS> struct ClassPrim;
template <> struct ClassPrim<CLS_POLY6> { typedef Prim<DCOL_POLYTOPE, 6> type; };
template <> struct ClassPrim<CLS_POLY8> { typedef Prim<DCOL_POLYTOPE, 8> type; };
template <> struct ClassPrim<CLS_POLYN> { typedef Prim<DCOL_POLYTOPE, 0> type; };
template <> struct ClassPrim<CLS_CAPSULE> { typedef Prim<DCOL_CAPSULE, 0> type; };
template <> struct ClassPrim<CLS_CYLINDER> { typedef Prim<DCOL_CYLINDER, 0> type; };
template <> struct ClassPrim<CLS_CONE> { typedef Prim<DCOL_CONE, 0> type; };
template <> struct ClassPrim<CLS_SPHERE> { typedef Prim<DCOL_SPHERE, 0> type; };
template <> struct ClassPrim<CLS_PGON5> { typedef Prim<DCOL_POLYGON, 5> type; };
template <> struct ClassPrim<CLS_PGONN> { typedef Prim<DCOL_POLYGON, 0> type; };
template <> struct ClassPrim<CLS_BOX> { typedef Prim<KIND_BOX, 6> type; };
template <> struct ClassPrim<CLS_ELLIPSOID> { typedef Prim<DCOL_ELLIPSOID, 0> type; };

/* faces exactly [I; -I] (create_rect_prism, misc_primitive_constructor.py:103-110) */
inline bool is_axis_box(const dcol_shape& s, const double* A)
{
    if (s.type != DCOL_POLYTOPE || s.n_faces != 6 || !A) return false;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 3; ++j)
            if (A[3 * (s.face_off + i) + j] != ((j == i % 3) ? (i < 3 ? 1.0 : -1.0) : 0.0)) return false;
    return true;
}

/* class of a shape record (A: the packed faces), or -1 if the record is malformed */
inline int shape_class(const dcol_shape& s, const double* A)
{
    switch (s.type) {
    case DCOL_POLYTOPE:
        if (s.n_faces < 1 || s.n_faces > DCOL_MAX_FACES) return -1;
        if (is_axis_box(s, A)) return CLS_BOX;
        return s.n_faces == 6 ? CLS_POLY6 : (s.n_faces == 8 ? CLS_POLY8 : CLS_POLYN);
    case DCOL_POLYGON:
        if (s.n_faces < 1 || s.n_faces > DCOL_MAX_FACES) return -1;
        return s.n_faces == 5 ? CLS_PGON5 : CLS_PGONN;
    case DCOL_CAPSULE: return CLS_CAPSULE;
    case DCOL_CYLINDER: return CLS_CYLINDER;
    case DCOL_CONE: return CLS_CONE;
    case DCOL_SPHERE: return CLS_SPHERE;
    case DCOL_ELLIPSOID: return (s.R > 0.0 && s.L > 0.0 && s.H > 0.0) ? CLS_ELLIPSOID : -1;
    default: return -1;
    }
}

inline bool class_has_extras(int cls)
{
    return cls == CLS_CAPSULE || cls == CLS_CYLINDER || cls == CLS_PGON5 || cls == CLS_PGONN;
}
/* combine_problem_matrices.py:58-67: the reference cannot assemble a pair in which both primitives
 * carry extra decision variables (np.vstack raises ValueError) -> DCOL_STATUS_UNSUPPORTED */
inline bool class_pair_supported(int c1, int c2) { return !(class_has_extras(c1) && class_has_extras(c2)); }

/* every pair of classes has a kernel; whether a both-extras pair is SOLVED is a run-time flag (DCOL_FIX_CASE4) */
template <int C1, int C2>
struct PairSupported {
    static constexpr bool value = true;
};

/* shape record + packed faces -> the uniform constant block of one specialisation */
template <int FMAX>
inline void fill_const(const dcol_shape& s, const double* A, const double* b, ShapeConst<FMAX>& c)
{
    c.R = s.R;
    c.L = s.L;
    c.H = s.H;
    c.tanb = tan(s.beta);
    c.ia[0] = s.R != 0.0 ? 1.0 / s.R : 0.0;
    c.ia[1] = s.L != 0.0 ? 1.0 / s.L : 0.0;
    c.ia[2] = s.H != 0.0 ? 1.0 / s.H : 0.0;
    for (int i = 0; i < 3; ++i) {
        c.r_off[i] = s.r_offset[i];
        for (int j = 0; j < 3; ++j) c.Q_off[i][j] = s.Q_offset[3 * i + j];
    }
    c.nf = s.n_faces;
    c.no_offset = 1;
    for (int i = 0; i < 3; ++i) {
        if (s.r_offset[i] != 0.0) c.no_offset = 0;
        for (int j = 0; j < 3; ++j)
            if (s.Q_offset[3 * i + j] != (i == j ? 1.0 : 0.0)) c.no_offset = 0;
    }
    for (int i = 0; i < 21; ++i) c.G0[i] = 0.0;
    for (int i = 0; i < (FMAX > 0 ? FMAX : 1); ++i) {
        c.A[i][0] = c.A[i][1] = c.A[i][2] = 0.0;
        c.b[i] = 0.0;
    }
    for (int i = 0; i < s.n_faces && i < FMAX; ++i) {
        for (int j = 0; j < 3; ++j) c.A[i][j] = A[3 * (s.face_off + i) + j];
        c.b[i] = b[s.face_off + i];
    }
}

/* the constant block of primitive class P: the shape's fields plus the constants derived from them */
template <class P>
inline void fill_const_prim(const dcol_shape& s, const double* A, const double* b, typename P::Const& c)
{
    fill_const(s, A, b, c);
    Solver<P, P>::template gram0<P>(c);
}

/* Runtime (c1, c2) -> compile-time pair of classes.  f must provide
 *   template <int C1, int C2> void operator()()        (called only for supported pairs) */
template <int C1, class F>
inline bool dispatch_class2(int c2, F& f)
{
#define DCOL_CASE2(C2)                                                   \
    case C2:                                                             \
        if constexpr (PairSupported<C1, C2>::value) {                    \
            f.template operator()<C1, C2>();                             \
            return true;                                                 \
        } else {                                                         \
            return false;                                                \
        }
    switch (c2) {
        DCOL_CASE2(CLS_POLY6)
        DCOL_CASE2(CLS_POLY8)
        DCOL_CASE2(CLS_POLYN)
        DCOL_CASE2(CLS_CAPSULE)
        DCOL_CASE2(CLS_CYLINDER)
        DCOL_CASE2(CLS_CONE)
        DCOL_CASE2(CLS_SPHERE)
        DCOL_CASE2(CLS_PGON5)
        DCOL_CASE2(CLS_PGONN)
        DCOL_CASE2(CLS_BOX)
        DCOL_CASE2(CLS_ELLIPSOID)
    default: return false;
    }
#undef DCOL_CASE2
}

template <class F>
inline bool dispatch_classes(int c1, int c2, F& f)
{
    switch (c1) {
    case CLS_POLY6: return dispatch_class2<CLS_POLY6>(c2, f);
    case CLS_POLY8: return dispatch_class2<CLS_POLY8>(c2, f);
    case CLS_POLYN: return dispatch_class2<CLS_POLYN>(c2, f);
    case CLS_CAPSULE: return dispatch_class2<CLS_CAPSULE>(c2, f);
    case CLS_CYLINDER: return dispatch_class2<CLS_CYLINDER>(c2, f);
    case CLS_CONE: return dispatch_class2<CLS_CONE>(c2, f);
    case CLS_SPHERE: return dispatch_class2<CLS_SPHERE>(c2, f);
    case CLS_PGON5: return dispatch_class2<CLS_PGON5>(c2, f);
    case CLS_PGONN: return dispatch_class2<CLS_PGONN>(c2, f);
    case CLS_BOX: return dispatch_class2<CLS_BOX>(c2, f);
    case CLS_ELLIPSOID: return dispatch_class2<CLS_ELLIPSOID>(c2, f);
    default: return false;
    }
}

} /* namespace dcol */
#endif /* DCOL_CLASSES_CUH_ */
