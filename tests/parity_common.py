"""Shared helpers of the parity tests: the same seeded inputs for the oracle and the CUDA path."""
import numpy as np

import hmgb200 as hmg
from oracle.mesh import Mesh as OMesh
from oracle.fem import build_local_diffusion_operators, build_local_mass_matrices, assemble_matrix
from oracle.interfaces import list_boundary_nodes_edges_faces, list_interior_nodes
from oracle.implicit import ImplicitFineGrid as OImplicit, ZeroDirichletConstraint
from oracle.operators import L2PlusDivAGrad
from oracle.multigrid import LevelState as OLevelState, BaseLevel as OBaseLevel


class Pair:
    """One problem, built twice: oracle objects (o_*) and the device context (g)."""

    def __init__(self, dim, c, levels, lam=1.0, field="checkerboard", seed=1, ordered=False, mesh=None, sigma=None,
                 device=0):
        if mesh is None:
            mesh, sigma = hmg.inputs.checkerboard_problem(dim, c, field=field, seed=seed, ordered=ordered)
        self.mesh, self.sigma, self.levels, self.lam, self.dim = mesh, sigma, levels, lam, mesh.dim
        self.obase = OMesh(mesh.nodes, mesh.elements)
        self.oimp = OImplicit(self.obase, levels)
        self.constraint = ZeroDirichletConstraint(*list_boundary_nodes_edges_faces(self.obase))
        diff = build_local_diffusion_operators(self.oimp.reference)
        mass = build_local_mass_matrices(self.oimp.reference)
        self.oops = [L2PlusDivAGrad(d, m, self.constraint, lam, sigma) for d, m in zip(diff, mass)]
        self.ostates = [OLevelState(self.oimp, l) for l in range(1, levels + 1)]
        self.g = hmg.ImplicitFineGrid(mesh, levels, sigma, lam=lam, device=device)
        self.rng = np.random.default_rng(seed + 100)

    def rand(self, level):
        return np.asfortranarray(self.rng.random((self.oimp.nf(level), self.mesh.nelements)))

    def obase_level(self):
        interior = list_interior_nodes(self.obase)
        A = assemble_matrix(self.obase, sigma=self.sigma, lam=self.lam)
        return OBaseLevel(A[interior][:, interior], self.obase.nnodes, interior), A[interior][:, interior], interior

    def sync_states_to_device(self, level):
        for name in ("x", "b", "r", "p", "Ap"):
            getattr(self.g.state(level), name).set(getattr(self.ostates[level - 1], name))

    def close(self):
        self.g.close()


def relerr(a, b):
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))
