"""smvs_b200 -- Blackwell-native Gauss-Newton depth refinement + SGM for the
SMVS pipeline (drop-in for flanggut/smvs' DepthOptimizer inner loop and
SGMStereo::run_sgm). The product is csrc/ (CUDA kernels + the C ABI declared
in include/smvs_b200.h); this package is the Python host-side mirror used by
the tests and the benchmark. It never imports anything from oracle/."""

__version__ = "0.1.0"
