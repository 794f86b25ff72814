// XLA FFI handlers over the C ABI of libqexxc.so (include/qexxc.h), for jax.ffi.
//
// NOT COMPILED OR TESTED IN THIS REPOSITORY'S ENVIRONMENT: JAX/jaxlib are not installable there
// (no network, not in the wheelhouse), so the XLA FFI headers ("xla/ffi/api/ffi.h", found under
// jax.ffi.include_dir()) are absent.  qex_b200/jax_ffi_shim.py compiles this file on demand when
// JAX is importable:  g++ -shared -fPIC -std=c++17 -I$(python -c 'import jax;print(jax.ffi.include_dir())')
//                     -I<repo>/include qexxc_jax_ffi.cc -L<repo>/qex_b200 -lqexxc -o libqexxc_jax.so
// Each handler is a thin adaptor: XLA buffers -> raw device pointers, XLA's stream -> cudaStream_t,
// return code -> ffi::Error.  The context handle travels as an int64 attribute.
#include <cstdint>

#include "qexxc.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using R64 = ffi::ResultBuffer<ffi::F64>;
typedef struct CUstream_st* cudaStream_t;

static ffi::Error Check(int rc) {
  if (rc == QEXXC_OK) return ffi::Error::Success();
  return ffi::Error(rc == QEXXC_ERR_ARG ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    qexxc_last_error());
}
static qexxc_ctx* Ctx(int64_t h) { return reinterpret_cast<qexxc_ctx*>(static_cast<intptr_t>(h)); }

// numint_legacy.py:122-348 (forward) --------------------------------------------------------------
static ffi::Error NrRksFwd(cudaStream_t s, int64_t ctx, int32_t xctype, int32_t hermi, F64 dm, F64 theta, R64 out,
                           R64 resid) {
  return Check(qexxc_nr_rks_fwd(Ctx(ctx), xctype, hermi, dm.typed_data(), theta.typed_data(), (long)theta.element_count(), out->typed_data(),
                                resid->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(QexxcNrRksFwd, NrRksFwd,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx")
                                  .Attr<int32_t>("xctype")
                                  .Attr<int32_t>("hermi")
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Ret<F64>()
                                  .Ret<F64>());

// reverse rule (trainer_legacy_no_jit.py:284) -------------------------------------------------------
static ffi::Error NrRksVjp(cudaStream_t s, int64_t ctx, int32_t xctype, int32_t hermi, F64 theta, F64 resid, F64 e_bar,
                           F64 v_bar, R64 bar) {
  return Check(qexxc_nr_rks_vjp(Ctx(ctx), xctype, hermi, theta.typed_data(), (long)theta.element_count(), resid.typed_data(), e_bar.typed_data(),
                                v_bar.typed_data(), bar->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(QexxcNrRksVjp, NrRksVjp,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx")
                                  .Attr<int32_t>("xctype")
                                  .Attr<int32_t>("hermi")
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Ret<F64>());

// eval_rho numint_legacy.py:351-397 and its reverse -------------------------------------------------
static ffi::Error EvalRho(cudaStream_t s, int64_t ctx, int32_t ncomp, int32_t hermi, F64 dm, R64 rho) {
  return Check(qexxc_eval_rho(Ctx(ctx), dm.typed_data(), ncomp, hermi, rho->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(QexxcEvalRho, EvalRho,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx")
                                  .Attr<int32_t>("ncomp")
                                  .Attr<int32_t>("hermi")
                                  .Arg<F64>()
                                  .Ret<F64>());
static ffi::Error EvalRhoVjp(cudaStream_t s, int64_t ctx, int32_t ncomp, int32_t hermi, F64 rho_bar, R64 dm_bar) {
  return Check(qexxc_eval_rho_vjp(Ctx(ctx), rho_bar.typed_data(), ncomp, hermi, dm_bar->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(QexxcEvalRhoVjp, EvalRhoVjp,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx")
                                  .Attr<int32_t>("ncomp")
                                  .Attr<int32_t>("hermi")
                                  .Arg<F64>()
                                  .Ret<F64>());

// network apply_fn (networks.py:43-75) and its reverse ------------------------------------------------
static ffi::Error ApplyFwd(cudaStream_t s, int64_t ctx, int64_t npts, F64 x, F64 theta, R64 y) {
  return Check(qexxc_apply_fn_fwd(Ctx(ctx), x.typed_data(), npts, theta.typed_data(), (long)theta.element_count(), y->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(QexxcApplyFwd, ApplyFwd,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx")
                                  .Attr<int64_t>("npts")
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Ret<F64>());
static ffi::Error ApplyVjp(cudaStream_t s, int64_t ctx, int64_t npts, F64 x, F64 theta, F64 y_bar, R64 x_bar,
                           R64 theta_bar) {
  return Check(qexxc_apply_fn_vjp(Ctx(ctx), x.typed_data(), npts, theta.typed_data(), (long)theta.element_count(), y_bar.typed_data(),
                                  x_bar->typed_data(), theta_bar->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(QexxcApplyVjp, ApplyVjp,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx")
                                  .Attr<int64_t>("npts")
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Ret<F64>()
                                  .Ret<F64>());
