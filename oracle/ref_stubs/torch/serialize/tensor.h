// Stand-in for <torch/serialize/tensor.h>, used ONLY to compile the reference's sptr CUDA kernel files where they lie
// (oracle/Makefile, target _ref/libsptr_ref.so).  Those .cu files include their .h, which declares pybind-side wrappers
// taking at::Tensor by value next to the extern "C" raw-pointer launchers this repo calls; a declaration only needs the
// name of the type.  Nothing of torch is compiled or linked into the reference library.
#pragma once
namespace at { class Tensor; }
