"""cadrays_b200 -- B200-native path tracer behind the OCCT calls CADRays makes (one hot path).

Modules (nothing is imported here, so `import cadrays_b200` has no side effects and does not load the library):
  view         V3d_View / Graphic3d_* host mirror over the C-ABI (include/cadrays_b200.h)
  scenes       procedural scenes of the BASELINE configs (C1..C5)
  tcl, ply     DRAW / Tcl-subset scene scripts, PLY meshes
  run, regress headless `script.tcl N` runner and the regression driver built on it
  imageio      PNG / HDR / PFM
  distributed  sample partition + all-reduce plumbing (torch.distributed)
  build, _ffi  nvcc build of libcadrays_b200.so, ctypes prototypes
"""
__version__ = "0.1.0"
