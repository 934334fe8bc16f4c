// glsl_vec_ops.inl — included INSIDE each tvec<T, N> specialisation of glsl_shim.h (test infrastructure).
// Operators and built-in functions are hidden friends: plain (non-template) functions found by argument-dependent
// lookup, so the implicit conversions GLSL allows (swizzle -> vector, int scalar -> float scalar) apply to their
// arguments the way they do in a GLSL compiler's closed overload set.
    typedef tvec V;
    typedef tvec<bool, dim> BV;
    explicit operator T() const { return d[0]; }   // GLSL: float(vec3) takes the first component (SURVEY Q2 relies on it)
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    friend V operator-(const V& a) { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(-a.d[i]); return r; }
    friend V operator+(const V& a) { return a; }

#define GLSL_VOP(OP) \
    friend V operator OP(const V& a, const V& b) { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a.d[i] OP b.d[i]); return r; } \
    friend V operator OP(const V& a, T b) { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a.d[i] OP b); return r; } \
    friend V operator OP(T a, const V& b) { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a OP b.d[i]); return r; } \
    V& operator OP##=(const V& b) { for (int i = 0; i < dim; ++i) d[i] = (T)(d[i] OP b.d[i]); return *this; } \
    V& operator OP##=(T b) { for (int i = 0; i < dim; ++i) d[i] = (T)(d[i] OP b); return *this; }
    GLSL_VOP(+) GLSL_VOP(-) GLSL_VOP(*) GLSL_VOP(/)
#undef GLSL_VOP
    // GLSL's implicit int -> float promotion for "integer vector (op) float scalar" (never the silent narrowing C++ would do)
#define GLSL_PROMO(OP) \
    friend tvec<float, dim> operator OP(const V& a, float b) requires (std::is_integral_v<T> && !std::is_same_v<T, bool>) \
    { tvec<float, dim> r; for (int i = 0; i < dim; ++i) r.d[i] = (float)a.d[i] OP b; return r; } \
    friend tvec<float, dim> operator OP(float a, const V& b) requires (std::is_integral_v<T> && !std::is_same_v<T, bool>) \
    { tvec<float, dim> r; for (int i = 0; i < dim; ++i) r.d[i] = a OP (float)b.d[i]; return r; }
    GLSL_PROMO(+) GLSL_PROMO(-) GLSL_PROMO(*) GLSL_PROMO(/)
#undef GLSL_PROMO
#define GLSL_IOP(OP) \
    friend V operator OP(const V& a, const V& b) requires std::is_integral_v<T> { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a.d[i] OP b.d[i]); return r; } \
    friend V operator OP(const V& a, T b) requires std::is_integral_v<T> { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a.d[i] OP b); return r; } \
    friend V operator OP(T a, const V& b) requires std::is_integral_v<T> { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a OP b.d[i]); return r; } \
    V& operator OP##=(const V& b) requires std::is_integral_v<T> { for (int i = 0; i < dim; ++i) d[i] = (T)(d[i] OP b.d[i]); return *this; } \
    V& operator OP##=(T b) requires std::is_integral_v<T> { for (int i = 0; i < dim; ++i) d[i] = (T)(d[i] OP b); return *this; }
    GLSL_IOP(%) GLSL_IOP(&) GLSL_IOP(|) GLSL_IOP(^)
#undef GLSL_IOP
    // shifts: the shift count may be int or uint whatever the vector's component type
    friend V operator<<(const V& a, int s) requires std::is_integral_v<T> { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a.d[i] << s); return r; }
    friend V operator>>(const V& a, int s) requires std::is_integral_v<T> { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a.d[i] >> s); return r; }
    friend V operator<<(const V& a, unsigned s) requires std::is_integral_v<T> { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a.d[i] << s); return r; }
    friend V operator>>(const V& a, unsigned s) requires std::is_integral_v<T> { V r; for (int i = 0; i < dim; ++i) r.d[i] = (T)(a.d[i] >> s); return r; }
    friend bool operator==(const V& a, const V& b) { for (int i = 0; i < dim; ++i) if (!(a.d[i] == b.d[i])) return false; return true; }
    friend bool operator!=(const V& a, const V& b) { return !(a == b); }

    // relational (GLSL 8.7)
#define GLSL_REL(NAME, OP) \
    friend BV NAME(const V& a, const V& b) { BV r; for (int i = 0; i < dim; ++i) r.d[i] = a.d[i] OP b.d[i]; return r; }
    GLSL_REL(lessThan, <) GLSL_REL(lessThanEqual, <=) GLSL_REL(greaterThan, >) GLSL_REL(greaterThanEqual, >=)
    GLSL_REL(equal, ==) GLSL_REL(notEqual, !=)
#undef GLSL_REL
    friend bool any(const V& a) requires std::is_same_v<T, bool> { for (int i = 0; i < dim; ++i) if (a.d[i]) return true; return false; }
    friend bool all(const V& a) requires std::is_same_v<T, bool> { for (int i = 0; i < dim; ++i) if (!a.d[i]) return false; return true; }

    // component-wise built-ins
#define GLSL_F1(NAME, EXPR) \
    friend V NAME(const V& a) requires std::is_floating_point_v<T> { V r; for (int i = 0; i < dim; ++i) { const T x = a.d[i]; r.d[i] = (EXPR); } return r; }
    GLSL_F1(floor, ::floorf(x)) GLSL_F1(ceil, ::ceilf(x)) GLSL_F1(fract, x - ::floorf(x)) GLSL_F1(sqrt, ::sqrtf(x))
    GLSL_F1(exp, ::expf(x)) GLSL_F1(exp2, ::exp2f(x)) GLSL_F1(log, ::logf(x)) GLSL_F1(log2, ::log2f(x))
    GLSL_F1(sin, ::sinf(x)) GLSL_F1(cos, ::cosf(x)) GLSL_F1(round, ::roundf(x))
#undef GLSL_F1
    friend V abs(const V& a) { V r; for (int i = 0; i < dim; ++i) r.d[i] = a.d[i] < (T)0 ? (T)(-a.d[i]) : a.d[i]; return r; }
    friend V pow(const V& a, const V& b) requires std::is_floating_point_v<T> { V r; for (int i = 0; i < dim; ++i) r.d[i] = a.d[i] < (T)0 ? (T)NAN : (T)::powf(a.d[i], b.d[i]); return r; }
    // floats: IEEE minNum / maxNum like the scalar versions in glsl_shim.h (x != x only for NaN)
    friend V min(const V& a, const V& b) { V r; for (int i = 0; i < dim; ++i) r.d[i] = (b.d[i] < a.d[i] || a.d[i] != a.d[i]) ? b.d[i] : a.d[i]; return r; }
    friend V max(const V& a, const V& b) { V r; for (int i = 0; i < dim; ++i) r.d[i] = (a.d[i] < b.d[i] || a.d[i] != a.d[i]) ? b.d[i] : a.d[i]; return r; }
    friend V min(const V& a, T b) { return min(a, V(b)); }
    friend V max(const V& a, T b) { return max(a, V(b)); }
    friend V clamp(const V& x, const V& lo, const V& hi) { return min(max(x, lo), hi); }
    friend V clamp(const V& x, T lo, T hi) { return min(max(x, V(lo)), V(hi)); }
    friend V mix(const V& x, const V& y, T a) requires std::is_floating_point_v<T> { return x * ((T)1 - a) + y * a; }
    friend V mix(const V& x, const V& y, const V& a) requires std::is_floating_point_v<T> { return x * (V((T)1) - a) + y * a; }
    // geometric (GLSL 8.5): dot summed left to right, normalize(v) = v / length(v)
    friend T dot(const V& a, const V& b) requires std::is_floating_point_v<T> { T s = a.d[0] * b.d[0]; for (int i = 1; i < dim; ++i) s = s + a.d[i] * b.d[i]; return s; }
    friend T length(const V& a) requires std::is_floating_point_v<T> { return ::sqrtf(dot(a, a)); }
    friend T distance(const V& a, const V& b) requires std::is_floating_point_v<T> { return length(a - b); }
    friend V normalize(const V& a) requires std::is_floating_point_v<T> { return a / length(a); }
    friend V reflect(const V& I, const V& N) requires std::is_floating_point_v<T> { return I - (T)2 * dot(N, I) * N; }
    friend V cross(const V& a, const V& b) requires (std::is_floating_point_v<T> && dim == 3)
    { return V(a.d[1] * b.d[2] - b.d[1] * a.d[2], a.d[2] * b.d[0] - b.d[2] * a.d[0], a.d[0] * b.d[1] - b.d[0] * a.d[1]); }

    // swizzles as member calls: `.xyz` in the shader text becomes `.xyz()` (build_ref.py rule R7). On a const vector the
    // call returns a value, otherwise an assignable proxy (swz<>) that converts to the value.
#define GLSL_MAXI2(A, B) ((A) > (B) ? (A) : (B))
#define GLSL_SW2(A, B, IA, IB) \
    tvec<T, 2> A##B() const requires (dim > GLSL_MAXI2(IA, IB)) { return tvec<T, 2>(d[IA], d[IB]); } \
    swz<tvec<T, 2>> A##B() requires (dim > GLSL_MAXI2(IA, IB)) { return swz<tvec<T, 2>>{ { &d[IA], &d[IB] } }; }
#define GLSL_SW3(A, B, C, IA, IB, IC) \
    tvec<T, 3> A##B##C() const requires (dim > GLSL_MAXI2(IA, GLSL_MAXI2(IB, IC))) { return tvec<T, 3>(d[IA], d[IB], d[IC]); } \
    swz<tvec<T, 3>> A##B##C() requires (dim > GLSL_MAXI2(IA, GLSL_MAXI2(IB, IC))) { return swz<tvec<T, 3>>{ { &d[IA], &d[IB], &d[IC] } }; }
#define GLSL_SW4(A, B, C, D, IA, IB, IC, ID) \
    tvec<T, 4> A##B##C##D() const requires (dim > GLSL_MAXI2(GLSL_MAXI2(IA, IB), GLSL_MAXI2(IC, ID))) { return tvec<T, 4>(d[IA], d[IB], d[IC], d[ID]); } \
    swz<tvec<T, 4>> A##B##C##D() requires (dim > GLSL_MAXI2(GLSL_MAXI2(IA, IB), GLSL_MAXI2(IC, ID))) { return swz<tvec<T, 4>>{ { &d[IA], &d[IB], &d[IC], &d[ID] } }; }

#define GLSL_SW2_B(A, IA, N0, N1, N2, N3) GLSL_SW2(A, N0, IA, 0) GLSL_SW2(A, N1, IA, 1) GLSL_SW2(A, N2, IA, 2) GLSL_SW2(A, N3, IA, 3)
#define GLSL_SW2_ALL(N0, N1, N2, N3) GLSL_SW2_B(N0, 0, N0, N1, N2, N3) GLSL_SW2_B(N1, 1, N0, N1, N2, N3) GLSL_SW2_B(N2, 2, N0, N1, N2, N3) GLSL_SW2_B(N3, 3, N0, N1, N2, N3)
#define GLSL_SW3_C(A, B, IA, IB, N0, N1, N2, N3) GLSL_SW3(A, B, N0, IA, IB, 0) GLSL_SW3(A, B, N1, IA, IB, 1) GLSL_SW3(A, B, N2, IA, IB, 2) GLSL_SW3(A, B, N3, IA, IB, 3)
#define GLSL_SW3_B(A, IA, N0, N1, N2, N3) GLSL_SW3_C(A, N0, IA, 0, N0, N1, N2, N3) GLSL_SW3_C(A, N1, IA, 1, N0, N1, N2, N3) GLSL_SW3_C(A, N2, IA, 2, N0, N1, N2, N3) GLSL_SW3_C(A, N3, IA, 3, N0, N1, N2, N3)
#define GLSL_SW3_ALL(N0, N1, N2, N3) GLSL_SW3_B(N0, 0, N0, N1, N2, N3) GLSL_SW3_B(N1, 1, N0, N1, N2, N3) GLSL_SW3_B(N2, 2, N0, N1, N2, N3) GLSL_SW3_B(N3, 3, N0, N1, N2, N3)
#define GLSL_SW4_D(A, B, C, IA, IB, IC, N0, N1, N2, N3) GLSL_SW4(A, B, C, N0, IA, IB, IC, 0) GLSL_SW4(A, B, C, N1, IA, IB, IC, 1) GLSL_SW4(A, B, C, N2, IA, IB, IC, 2) GLSL_SW4(A, B, C, N3, IA, IB, IC, 3)
#define GLSL_SW4_C(A, B, IA, IB, N0, N1, N2, N3) GLSL_SW4_D(A, B, N0, IA, IB, 0, N0, N1, N2, N3) GLSL_SW4_D(A, B, N1, IA, IB, 1, N0, N1, N2, N3) GLSL_SW4_D(A, B, N2, IA, IB, 2, N0, N1, N2, N3) GLSL_SW4_D(A, B, N3, IA, IB, 3, N0, N1, N2, N3)
#define GLSL_SW4_B(A, IA, N0, N1, N2, N3) GLSL_SW4_C(A, N0, IA, 0, N0, N1, N2, N3) GLSL_SW4_C(A, N1, IA, 1, N0, N1, N2, N3) GLSL_SW4_C(A, N2, IA, 2, N0, N1, N2, N3) GLSL_SW4_C(A, N3, IA, 3, N0, N1, N2, N3)
#define GLSL_SW4_ALL(N0, N1, N2, N3) GLSL_SW4_B(N0, 0, N0, N1, N2, N3) GLSL_SW4_B(N1, 1, N0, N1, N2, N3) GLSL_SW4_B(N2, 2, N0, N1, N2, N3) GLSL_SW4_B(N3, 3, N0, N1, N2, N3)
    GLSL_SW2_ALL(x, y, z, w) GLSL_SW3_ALL(x, y, z, w) GLSL_SW4_ALL(x, y, z, w)
    GLSL_SW2_ALL(r, g, b, a) GLSL_SW3_ALL(r, g, b, a) GLSL_SW4_ALL(r, g, b, a)
#undef GLSL_SW2_ALL
#undef GLSL_SW3_ALL
#undef GLSL_SW4_ALL
#undef GLSL_SW2_B
#undef GLSL_SW3_B
#undef GLSL_SW3_C
#undef GLSL_SW4_B
#undef GLSL_SW4_C
#undef GLSL_SW4_D
#undef GLSL_SW2
#undef GLSL_SW3
#undef GLSL_SW4
#undef GLSL_MAXI2
