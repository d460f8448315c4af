// torch_op.cpp -- TorchScript-visible operators over the C ABI of libgnnis.so ("gradient contract").
//
// The reference hands OpenMM a TorchScript file (`openmmtorch.TorchForce(file)`, Simulation/helper_functions.py:777-785);
// TorchForce calls the scripted `forward(positions)` from C++ and either runs `energy.backward()` and reads
// `positions.grad`, or takes `(energy, forces)` when `setOutputsForces(True)`.  A Python `autograd.Function` cannot be
// scripted, so the two calls are registered here as custom operators that a scripted module can reach:
//
//   torch.ops.gnnis.energy_force(pos, handle, n_rep, flags) -> (E[n_rep], F like pos)      no autograd
//   torch.ops.gnnis.energy(pos, handle, n_rep, flags)       -> E[n_rep]                    d/dpos = -F
//
// `handle` is the gnnis_handle of a live Engine in the same process (an int64 attribute of the scripted module).
// Only torch types cross this file; everything below it is the plain C ABI of include/gnnis.h.
#include <torch/script.h>
#include <torch/autograd.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <tuple>

#include "gnnis.h"

namespace {

std::tuple<torch::Tensor, torch::Tensor> energy_force_impl(const torch::Tensor& pos, int64_t handle, int64_t n_rep,
                                                           int64_t flags, bool want_forces) {
  TORCH_CHECK(pos.is_cuda(), "gnnis: positions must be a CUDA tensor");
  TORCH_CHECK(pos.scalar_type() == torch::kFloat32, "gnnis: positions must be float32");
  TORCH_CHECK(handle != 0 && n_rep > 0, "gnnis: invalid handle / replica count");
  const c10::cuda::CUDAGuard guard(pos.device());
  torch::Tensor p = pos.contiguous();
  torch::Tensor e = torch::empty({n_rep}, p.options());
  torch::Tensor f = want_forces ? torch::empty_like(p) : torch::Tensor();
  gnnis_handle h = reinterpret_cast<gnnis_handle>(handle);
  uint32_t fl = static_cast<uint32_t>(flags) | (want_forces ? 0u : GNNIS_FLAG_NO_FORCES);
  int rc = gnnis_energy_force(h, p.data_ptr<float>(), e.data_ptr<float>(), want_forces ? f.data_ptr<float>() : nullptr, fl,
                              c10::cuda::getCurrentCUDAStream(pos.device().index()).stream());
  TORCH_CHECK(rc == 0, "gnnis_energy_force failed (", rc, "): ", gnnis_last_error(h));
  return {e, f};
}

std::tuple<torch::Tensor, torch::Tensor> energy_force(const torch::Tensor& pos, int64_t handle, int64_t n_rep,
                                                      int64_t flags) {
  return energy_force_impl(pos, handle, n_rep, flags, true);
}

// E[n_rep] with an autograd edge to the positions: the forward stashes F, the backward returns -grad_E[rep] * F
struct EnergyFn : public torch::autograd::Function<EnergyFn> {
  static torch::Tensor forward(torch::autograd::AutogradContext* ctx, const torch::Tensor& pos, int64_t handle,
                               int64_t n_rep, int64_t flags) {
    at::AutoDispatchBelowADInplaceOrView g;
    auto ef = energy_force_impl(pos, handle, n_rep, flags, true);
    ctx->save_for_backward({std::get<1>(ef)});
    ctx->saved_data["n_rep"] = n_rep;
    return std::get<0>(ef);
  }
  static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx,
                                               torch::autograd::tensor_list grad_out) {
    const torch::Tensor f = ctx->get_saved_variables()[0];
    const int64_t n_rep = ctx->saved_data["n_rep"].toInt();
    // forces are per atom of replica r: broadcast grad_E[r] over the atoms and components of that replica
    torch::Tensor g = grad_out[0].contiguous().view({n_rep, 1});
    torch::Tensor gp = (-(f.view({n_rep, -1}) * g)).view_as(f);
    return {gp, torch::Tensor(), torch::Tensor(), torch::Tensor()};
  }
};

torch::Tensor energy_autograd(const torch::Tensor& pos, int64_t handle, int64_t n_rep, int64_t flags) {
  return EnergyFn::apply(pos, handle, n_rep, flags);
}

torch::Tensor energy_cuda(const torch::Tensor& pos, int64_t handle, int64_t n_rep, int64_t flags) {
  return std::get<0>(energy_force_impl(pos, handle, n_rep, flags, false));
}

}  // namespace

TORCH_LIBRARY(gnnis, m) {
  m.def("energy_force(Tensor pos, int handle, int n_rep, int flags) -> (Tensor, Tensor)");
  m.def("energy(Tensor pos, int handle, int n_rep, int flags) -> Tensor");
}
TORCH_LIBRARY_IMPL(gnnis, CUDA, m) {
  m.impl("energy_force", energy_force);
  m.impl("energy", energy_cuda);
}
TORCH_LIBRARY_IMPL(gnnis, Autograd, m) { m.impl("energy", energy_autograd); }
