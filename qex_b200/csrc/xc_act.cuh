// Activation functions of the reference's MLPs with first and second derivatives
// (qedft/models/classical/classical_models.py:39-49 ACTIVATION_MAP; flax nn.gelu / nn.swish of
// trainer_legacy_no_jit.py:96-107).  stax.Gelu / nn.gelu are the tanh-approximate GELU.
// Accurate libm-grade device functions only (no __expf-style intrinsics): the float32 path
// has to hold 1e-5 relative against the float64 reference.
#pragma once
#include "../../include/qexxc.h"

namespace qexxc {

template <typename T>
struct Vec2;
template <>
struct Vec2<double> {
    typedef double2 type;
    __device__ static __forceinline__ double2 make(double a, double b) { return make_double2(a, b); }
};
template <>
struct Vec2<float> {
    typedef float2 type;
    __device__ static __forceinline__ float2 make(float a, float b) { return make_float2(a, b); }
};

__device__ __forceinline__ double qx_tanh(double x) { return tanh(x); }
__device__ __forceinline__ float qx_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double qx_exp(double x) { return exp(x); }
__device__ __forceinline__ float qx_exp(float x) { return expf(x); }
__device__ __forceinline__ double qx_log1p(double x) { return log1p(x); }
__device__ __forceinline__ float qx_log1p(float x) { return log1pf(x); }

// s0 = sigma(z), s1 = sigma'(z), s2 = sigma''(z)
template <typename T>
__device__ __forceinline__ void act_d012(int act, T z, T& s0, T& s1, T& s2) {
    switch (act) {
        case QEXXC_ACT_TANH: {
            const T t = qx_tanh(z);
            s0 = t;
            s1 = (T)1 - t * t;
            s2 = (T)-2 * t * s1;
            break;
        }
        case QEXXC_ACT_SIGMOID: {
            const T s = (T)1 / ((T)1 + qx_exp(-z));
            s0 = s;
            s1 = s * ((T)1 - s);
            s2 = s1 * ((T)1 - (T)2 * s);
            break;
        }
        case QEXXC_ACT_SOFTPLUS: {
            const T s = (T)1 / ((T)1 + qx_exp(-z));
            s0 = (z > (T)0 ? z : (T)0) + qx_log1p(qx_exp(z > (T)0 ? -z : z));
            s1 = s;
            s2 = s * ((T)1 - s);
            break;
        }
        case QEXXC_ACT_RELU: {
            const T p = z > (T)0 ? (T)1 : (T)0;
            s0 = z * p;
            s1 = p;
            s2 = (T)0;
            break;
        }
        case QEXXC_ACT_LEAKY_RELU: {
            const T p = z >= (T)0 ? (T)1 : (T)0.01;
            s0 = z * p;
            s1 = p;
            s2 = (T)0;
            break;
        }
        case QEXXC_ACT_ELU: {
            const T e = qx_exp(z < (T)0 ? z : (T)0);
            const bool pos = z > (T)0;
            s0 = pos ? z : e - (T)1;
            s1 = pos ? (T)1 : e;
            s2 = pos ? (T)0 : e;
            break;
        }
        case QEXXC_ACT_SELU: {
            const T lam = (T)1.0507009873554804934193349852946, al = (T)1.6732632423543772848170429916717;
            const T e = qx_exp(z < (T)0 ? z : (T)0);
            const bool pos = z > (T)0;
            s0 = lam * (pos ? z : al * (e - (T)1));
            s1 = lam * (pos ? (T)1 : al * e);
            s2 = lam * (pos ? (T)0 : al * e);
            break;
        }
        case QEXXC_ACT_GELU: {
            const T k = (T)0.7978845608028654, c = (T)0.044715;
            const T u = k * (z + c * z * z * z);
            const T u1 = k * ((T)1 + (T)3 * c * z * z);
            const T u2 = k * (T)6 * c * z;
            const T t = qx_tanh(u);
            const T s = (T)1 - t * t;
            s0 = (T)0.5 * z * ((T)1 + t);
            s1 = (T)0.5 * ((T)1 + t) + (T)0.5 * z * s * u1;
            s2 = s * u1 + (T)0.5 * z * ((T)-2 * t * s * u1 * u1 + s * u2);
            break;
        }
        default: {  // QEXXC_ACT_SWISH
            const T s = (T)1 / ((T)1 + qx_exp(-z));
            const T ds = s * ((T)1 - s);
            s0 = z * s;
            s1 = s + z * ds;
            s2 = (T)2 * ds + z * ds * ((T)1 - (T)2 * s);
            break;
        }
    }
}

}  // namespace qexxc
