# baseline/julia_cpu_baseline.jl -- times the UNMODIFIED reference (haampie/Homogenization.jl) on the host cores.
#
# Julia is not installed in the image this repository is developed and benchmarked in, so bench.py's reference arm
# runs the threaded C restatement of the same algorithm (oracle/c/hmg_ref_cpu.c) instead.  Wherever Julia >= 1.0 and
# the reference package are available, this script gives the real number for the same shapes:
#
#     JULIA_NUM_THREADS=$(nproc) julia -O3 baseline/julia_cpu_baseline.jl 3 14 4        # dim, cells per side, refinements
#
# It times (i) the global product  fill!; mul!; apply_constraint!; broadcast_interfaces!  of src/multigrid.jl:58-61 and
# (ii) vcycle!(..., 3) (src/multigrid.jl:73-119) on a checkerboard field and prints one JSON line in bench.py's units
# (stored finest-level DOFs per second).  Inputs use Julia's RNG: throughput does not depend on the draw.
using Homogenization, LinearAlgebra, SparseArrays, StaticArrays, Random
using Homogenization: hypercube, Tri64, Tet64, ImplicitFineGrid, list_boundary_nodes_edges_faces, ZeroDirichletConstraint,
                      build_local_diffusion_operators, build_local_mass_matrices, L2PlusDivAGrad, LevelState, BaseLevel,
                      refined_mesh, nnodes, nelements, list_interior_nodes, assemble_checkerboard, generate_conductivity,
                      conductivity_per_element, broadcast_interfaces!, apply_constraint!, local_rhs!, vcycle!, base_mesh

function main(dim::Int, c::Int, refinements::Int; steps = 3)
    Random.seed!(1)
    base = hypercube(dim == 2 ? Tri64 : Tet64, c, origin = ntuple(_ -> -c / 2, dim))
    σ_cells = generate_conductivity(base, c)
    cond = conductivity_per_element(base, σ_cells, SVector{dim,Float64}(ntuple(_ -> c / 2 + 1.0, dim)))
    grids = refinements + 1
    implicit = ImplicitFineGrid(base, grids)
    constraint = ZeroDirichletConstraint(list_boundary_nodes_edges_faces(base)...)
    diff = build_local_diffusion_operators(implicit.reference)
    mass = build_local_mass_matrices(implicit.reference)
    ops = [L2PlusDivAGrad(d, m, constraint, 1.0, cond) for (d, m) in zip(diff, mass)]
    states = [LevelState(nelements(base), nnodes(refined_mesh(implicit, l)), Float64) for l = 1 : grids]
    top = states[end]
    rand!(top.x); broadcast_interfaces!(top.x, implicit, grids); apply_constraint!(top.x, grids, constraint, implicit)
    local_rhs!(top.b, implicit)
    interior = list_interior_nodes(base)
    F = cholesky(assemble_checkerboard(base, cond, 1.0)[interior, interior])
    base_level = BaseLevel(Float64, F, nnodes(base), interior)
    dofs = length(top.x)

    copyto!(top.p, top.x)
    product!() = (fill!(top.Ap, 0.0); mul!(1.0, base_mesh(implicit), ops[end], top.p, top.Ap);
                  apply_constraint!(top.Ap, grids, constraint, implicit); broadcast_interfaces!(top.Ap, implicit, grids))
    product!()
    t_ax = @elapsed for _ = 1 : steps; product!(); end
    vcycle!(implicit, base_level, ops, states, grids, 3)
    t_v = @elapsed for _ = 1 : steps; vcycle!(implicit, base_level, ops, states, grids, 3); end
    println("{\"impl\": \"reference (Julia)\", \"threads\": $(Threads.nthreads()), \"dim\": $dim, \"cells_per_side\": $c, ",
            "\"refinements\": $refinements, \"stored_dofs\": $dofs, \"ax_gdofs\": $(dofs * steps / t_ax / 1e9), ",
            "\"vcycle_gdofs\": $(dofs * steps / t_v / 1e9)}")
end

main(parse(Int, ARGS[1]), parse(Int, ARGS[2]), parse(Int, ARGS[3]))
