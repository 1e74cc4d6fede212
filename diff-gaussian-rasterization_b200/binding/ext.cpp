// ext.cpp — pybind11 module `_C` exporting the same three functions as the reference
// (diff-gaussian-rasterization-{light,full}/ext.cpp:15-19).
#include <torch/extension.h>

#include "rasterize_points.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
  m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
  // extension: the same call + (want_colors_grad, want_cov3D_grad), see rasterize_points.cpp
  m.def("rasterize_gaussians_backward_select", &RasterizeGaussiansBackwardSelect);
  m.def("mark_visible", &markVisible);
  // extension (not in the reference): flat scene-gradient arena for view-level data parallelism
  m.def("set_grad_arena", &setGradArena, pybind11::arg("arena"), pybind11::arg("factorized_sh") = false,
        pybind11::arg("early_masked_color") = false);
  m.def("wait_masked_color", &waitMaskedColor, pybind11::arg("stream"));
  m.def("arm_grad_arena", &armGradArena);
  m.def("sh_grad_from_views", &shGradFromViews);
  // extensions for the NVLS exchange of dp.py (mode "nvls"): P2P-reading SH rebuild, in-switch slice all-reduce
  m.def("sh_grad_from_view_ptrs", &shGradFromViewPtrs);
  m.def("nvls_allreduce_slice", &nvlsAllreduceSlice, pybind11::arg("multicast_ptr"), pybind11::arg("offset_floats"),
        pybind11::arg("count_floats"), pybind11::arg("rank"), pybind11::arg("world"), pybind11::arg("max_blocks") = 0);
  m.def("p2p_gather", &p2pGather, pybind11::arg("src_ptrs"), pybind11::arg("count_floats"), pybind11::arg("dst"),
        pybind11::arg("max_blocks") = 0);
  m.def("p2p_allreduce_slice", &p2pAllreduceSlice, pybind11::arg("replica_ptrs"), pybind11::arg("offset_floats"),
        pybind11::arg("count_floats"), pybind11::arg("rank"), pybind11::arg("max_blocks") = 0);
  // extension (not in the reference): in-kernel densification statistics
  // extension (not in the reference): RGB-D L1 loss + cotangents in one pass over the rendered images
  m.def("rgbd_l1_loss", &rgbdL1Loss);
  m.def("set_densify_stats", &setDensifyStats, pybind11::arg("grad_accum"), pybind11::arg("denom"),
        pybind11::arg("max_radii2D"));
}
