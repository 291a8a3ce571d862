"""CPU oracle for the Homogenization.jl hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the reference's algorithm (the
implicit fine grid, the matrix-free local operators, the interface sums and the
geometric multigrid V-cycle, plus the few driver pieces needed for the sigma
history).  Every function cites the reference file:line it follows (paths are
relative to the reference repository root).

It is the CHECKER for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.  The
product (``homogenization.jl_b200``) never imports, links or executes anything
from this directory and fails loudly when its CUDA library is missing.

Parity status
-------------
* The reference is pure Julia and Julia is not installed in this image, so the
  reference cannot be executed here and ``oracle/_ref`` does not exist
  ("unbuildable": there is no C/C++ source to compile).
* A*x is PINNED by the reference's own known-answer test
  (test/test_operator.jl:9-73: implicit product == assembled-matrix product
  within 20 eps), restated in tests/test_oracle_reference_tests.py together with
  test/implicit_grid.jl, test/interpolation.jl, test/refined_reference_element.jl,
  test/sparse_cell_to_element.jl, test/generated_grids.jl, test/tricks.jl,
  test/counting_sort.jl and test/bitonic.jl.
* The V-cycle residual history and the homogenized coefficient are PARITY
  UNPINNED by the reference: no reference test executes vcycle!/cholesky and all
  of its randomness is unseeded.  They are pinned only by this restatement,
  whose multigrid is validated by convergence to the direct solution of the
  explicitly assembled fine problem.

Index convention: everything in this package is 0-based (numpy); the reference
is 1-based.  Orderings (which is what matters) are identical.
"""

from .mesh import Mesh, edge_graph, refine_uniformly, hypercube, sort_element_nodes  # noqa: F401
