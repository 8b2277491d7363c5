// Stand-in for pybind11/eigen.h when the reference's python/bindings/common.hpp is compiled here: Eigen is
// not installed in this image and no binding on the MCTS path passes an Eigen type.
#pragma once
