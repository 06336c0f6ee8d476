"""Refiner MLP (second-stage pose refinement) on the libepb.so kernels: model, loop, data stand-ins."""
