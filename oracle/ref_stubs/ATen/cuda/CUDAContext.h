// Stand-in for <ATen/cuda/CUDAContext.h> (see torch/serialize/tensor.h in this directory): the sptr kernel files use nothing from it.
#pragma once
