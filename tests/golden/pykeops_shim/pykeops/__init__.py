"""TEST INFRASTRUCTURE — a dense, pure-torch stand-in for the third-party ``pykeops`` package.

The reference (jeanfeydy/geomloss @ 00e493f) delegates its "online", "multiscale" and grid reductions to
pykeops (un-vendored, un-pinned: pyproject.toml:33-36), which JIT-compiles CUDA and cannot be installed in
the build container.  This shim implements exactly the pykeops surface the reference calls, with dense
torch tensors on the CPU (fp32 or fp64, autograd-capable), so that the UNMODIFIED reference code under
/root/reference/src can run its multiscale / online / grid / barycenter paths here and generate golden
vectors (tests/golden/make_golden_multiscale.py, make_golden_images.py).

Call sites served (reference file:line):
  generic_logsumexp(+ranges)   _legacy/sinkhorn_samples.py:325-333, :433-441, :448-450
  LazyTensor subset            _legacy/sinkhorn_samples.py:273-288, _legacy/utils.py:29-38, :152-161, :247-259,
                               _legacy/kernel_samples.py:62-82 (K.ranges = ...), :128-137 (K @ v, K.t())
  grid_cluster, cluster_ranges_centroids, sort_clusters, from_matrix, swap_axes
                               _legacy/sinkhorn_samples.py:477-479, :515, :529, _legacy/kernel_samples.py:224-256

Semantics follow pykeops' documented public API (v2.x): see each function's docstring.  Nothing here is
product code; the product never imports it.
"""
__version__ = "shim-2.x"
