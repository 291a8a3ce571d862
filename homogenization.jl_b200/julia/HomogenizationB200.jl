# HomogenizationB200.jl -- the reference-side binding of libhmg_b200.so (include/hmg.h).
#
# This file is what a maintainer of haampie/Homogenization.jl adds to route the hot path
# (matrix-free A*x on the implicit fine grid + the multigrid V-cycle) to a B200 through `ccall`.
# It adds METHODS to the reference's own generic functions (mul!, broadcast_interfaces!,
# apply_constraint!, zero_out_all_but_one!, local_residual!, restrict_to!, interpolate_and_sum_to!,
# smoothing_steps!, vcycle!) for a device matrix type, so `checkerboard_homogenization` runs
# unchanged once its LevelStates are built with `DeviceLevelStates`.
#
# NOTE: Julia is not installed in the image this repository is developed in; this shim is written
# against the reference's signatures (cited below, paths relative to the reference repository) and
# the C ABI that the Python ctypes binding (homogenization.jl_b200/_lib.py) exercises on every test
# run, but it has not itself been executed.
module HomogenizationB200

using Homogenization
using Homogenization: Mesh, ImplicitFineGrid, LevelState, BaseLevel, L2PlusDivAGrad, SimpleDiffusion,
                      ZeroDirichletConstraint, nelements, nnodes, refined_mesh, nlevels, construct_full_grid
using WriteVTK: vtk_grid, vtk_point_data
import Homogenization: export_unknown, local_rhs!, broadcast_interfaces!, apply_constraint!, zero_out_all_but_one!,
                       local_residual!, restrict_to!, interpolate_and_sum_to!, smoothing_steps!,
                       vcycle!, rhs_aξ∇v!, integrate_first_term, integrate_terms, integrate_area, next_rhs!
import LinearAlgebra: mul!
using SparseArrays: SparseMatrixCSC
using StaticArrays: SVector

const libhmg = get(ENV, "HMG_B200_LIB", joinpath(@__DIR__, "..", "libhmg_b200.so"))

# state vector ids of include/hmg.h (enum hmg_vec)
const HMG_X, HMG_B, HMG_R, HMG_P, HMG_AP, HMG_V, HMG_W = Cint.(0:6)

struct HmgError <: Exception
    msg::String
end

@inline function check(status::Cint)
    status == 0 && return nothing
    throw(HmgError(unsafe_string(ccall((:hmg_last_error, libhmg), Cstring, ()))))
end

"""
One context = ImplicitFineGrid + ZeroDirichletConstraint + L2PlusDivAGrad on every level + the
LevelStates, resident on one GPU (hmg_create).  Freed by a finalizer (hmg_destroy).
"""
mutable struct DeviceGrid
    ctx::Ptr{Cvoid}
    implicit::ImplicitFineGrid
    σs_hash::UInt          # hash of the coefficient vector the context holds (see sync_operator!)
    constraint::Any        # the ZeroDirichletConstraint the context was built for: derived from `implicit.base` at creation
    DeviceGrid(ctx::Ptr{Cvoid}, implicit::ImplicitFineGrid) = new(ctx, implicit, UInt(0), nothing)
    function DeviceGrid(implicit::ImplicitFineGrid{dim}, σs::Vector{SVector{dim,Float64}}, λ::Float64; device::Integer = 0) where {dim}
        base = implicit.base
        ctx = Ref{Ptr{Cvoid}}(C_NULL)
        nodes = reinterpret(Float64, base.nodes)          # dim x nn, column-major
        elems = reinterpret(Int64, base.elements)         # (dim+1) x ne, 1-based, sorted per element
        sig = reinterpret(Float64, σs)                    # dim x ne
        GC.@preserve nodes elems sig begin
            check(ccall((:hmg_create, libhmg), Cint,
                (Cint, Cint, Int64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Float64, Cint, Ref{Ptr{Cvoid}}),
                dim, nlevels(implicit), nelements(base), nnodes(base), nodes, elems, sig, λ, device, ctx))
        end
        g = new(ctx[], implicit, hash(σs), nothing)
        finalizer(x -> (x.ctx == C_NULL || ccall((:hmg_destroy, libhmg), Cint, (Ptr{Cvoid},), x.ctx); x.ctx = C_NULL), g)
        return g
    end
end

"""
An Nf(level) x Ne matrix living on the GPU: the `Tv <: AbstractMatrix{T}` of
`LevelState{T,Tv}` (src/multigrid.jl:7).  Scalar indexing goes through a download and is meant
for debugging only.
"""
struct DeviceMatrix <: AbstractMatrix{Float64}
    grid::DeviceGrid
    level::Cint
    which::Cint
end

Base.size(A::DeviceMatrix) = (Int(ccall((:hmg_nf, libhmg), Int64, (Ptr{Cvoid}, Cint), A.grid.ctx, A.level)),
                              Int(ccall((:hmg_ne_local, libhmg), Int64, (Ptr{Cvoid},), A.grid.ctx)))

function Base.copyto!(dst::DeviceMatrix, src::Matrix{Float64})
    @assert size(dst) == size(src)
    GC.@preserve src check(ccall((:hmg_upload, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int64),
                                 dst.grid.ctx, dst.level, dst.which, src, size(src, 1)))
    dst
end
function Base.copyto!(dst::Matrix{Float64}, src::DeviceMatrix)
    @assert size(dst) == size(src)
    GC.@preserve dst check(ccall((:hmg_download, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int64),
                                 src.grid.ctx, src.level, src.which, dst, size(dst, 1)))
    dst
end
function Base.copyto!(dst::DeviceMatrix, src::DeviceMatrix)
    @assert dst.level == src.level
    check(ccall((:hmg_copy, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint, Cint), dst.grid.ctx, dst.level, dst.which, src.which))
    dst
end
Base.Array(A::DeviceMatrix) = copyto!(Matrix{Float64}(undef, size(A)...), A)
Base.getindex(A::DeviceMatrix, i::Int, j::Int) = Array(A)[i, j]      # debugging only
Base.fill!(A::DeviceMatrix, v::Real) = (check(ccall((:hmg_fill, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint, Float64),
                                                    A.grid.ctx, A.level, A.which, Float64(v))); A)

"""
    DeviceLevelStates(grid) -> Vector{LevelState{Float64,DeviceMatrix}}

Replaces `LevelState(nelements(base), nnodes(mesh), Float64)` of
src/examples/homogenized_coefficients.jl:237-240.
"""
DeviceLevelStates(g::DeviceGrid) = map(1:nlevels(g.implicit)) do k
    LevelState{Float64,DeviceMatrix}((DeviceMatrix(g, k, w) for w in (HMG_X, HMG_B, HMG_R, HMG_P, HMG_AP))...)
end

# ---- device methods of the reference's generic functions -------------------------------------

"""
    sync_operator!(grid, A::L2PlusDivAGrad)

The context holds its own copy of the operator data.  Every device method that applies `A` forwards what the
reference's operator struct may have had mutated since (src/build_local_operators.jl:26-32 is a `mutable struct`):
λ always (a changed λ re-factorises an internally assembled coarse matrix on the next V-cycle), σs when its contents
changed (`hmg_set_sigma`).  The Dirichlet constraint is derived from the base mesh when the context is created; an
operator whose `constraint` is another object than the first one seen (the reference's driver builds a new one after
every domain shrink, src/examples/homogenized_coefficients.jl:331-333) needs a new `DeviceGrid` on the shrunk mesh --
that is an error here, not a silent mismatch.
"""
function sync_operator!(g::DeviceGrid, A::L2PlusDivAGrad)
    check(ccall((:hmg_set_lambda, libhmg), Cint, (Ptr{Cvoid}, Float64), g.ctx, A.λ))
    h = hash(A.σs)
    if h != g.σs_hash
        @assert length(A.σs) == nelements(g.implicit.base) "σs does not belong to the mesh of this DeviceGrid"
        sig = reinterpret(Float64, A.σs)
        GC.@preserve sig check(ccall((:hmg_set_sigma, libhmg), Cint, (Ptr{Cvoid}, Ptr{Float64}), g.ctx, sig))
        g.σs_hash = h
    end
    if g.constraint === nothing
        g.constraint = A.constraint
    elseif g.constraint !== A.constraint
        throw(HmgError("the operator carries another constraint than the one this DeviceGrid was used with: " *
                       "a new constraint (domain shrink) requires a new DeviceGrid on the shrunk mesh"))
    end
    nothing
end

# mul!(α, base, A, x, y): y ← αAx + y   (src/apply_local_operators.jl:85-91)
function mul!(α::Float64, base::Mesh, A::L2PlusDivAGrad, x::DeviceMatrix, y::DeviceMatrix)
    sync_operator!(x.grid, A)
    check(ccall((:hmg_mul, libhmg), Cint, (Ptr{Cvoid}, Cint, Float64, Cint, Cint), x.grid.ctx, x.level, α, x.which, y.which))
    y
end

# broadcast_interfaces!(x, implicit, level)   (src/implicit_fine_grid.jl:209)
broadcast_interfaces!(x::DeviceMatrix, implicit::ImplicitFineGrid, level::Int) =
    (check(ccall((:hmg_broadcast_interfaces, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint), x.grid.ctx, level, x.which)); x)

# apply_constraint!(x, level, z, implicit)   (src/implicit_fine_grid.jl:94)
apply_constraint!(x::DeviceMatrix, level::Int, z::ZeroDirichletConstraint, implicit::ImplicitFineGrid) =
    (check(ccall((:hmg_apply_constraint, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint), x.grid.ctx, level, x.which)); x)

# zero_out_all_but_one!(x, implicit, level)   (src/implicit_fine_grid.jl:334)
zero_out_all_but_one!(x::DeviceMatrix, implicit::ImplicitFineGrid, level::Int) =
    (check(ccall((:hmg_zero_out_all_but_one, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint), x.grid.ctx, level, x.which)); x)

# local_residual!(implicit, A, curr, k)   (src/apply_local_operators.jl:18-27)
function local_residual!(implicit::ImplicitFineGrid, A::L2PlusDivAGrad, curr::LevelState{Float64,DeviceMatrix}, k::Int)
    sync_operator!(curr.x.grid, A)
    check(ccall((:hmg_local_residual, libhmg), Cint, (Ptr{Cvoid}, Cint), curr.x.grid.ctx, k))
end

# restrict_to!(next.b, P, curr.r) / interpolate_and_sum_to!(curr.x, P, next.x)   (src/interpolation.jl:52-74)
restrict_to!(y::DeviceMatrix, P::SparseMatrixCSC, x::DeviceMatrix) =
    check(ccall((:hmg_restrict, libhmg), Cint, (Ptr{Cvoid}, Cint), x.grid.ctx, x.level))
interpolate_and_sum_to!(y::DeviceMatrix, P::SparseMatrixCSC, x::DeviceMatrix) =
    check(ccall((:hmg_interpolate_add, libhmg), Cint, (Ptr{Cvoid}, Cint), y.grid.ctx, y.level))

# smoothing_steps!(steps, implicit, ops, curr, k)   (src/multigrid.jl:46-71)
function smoothing_steps!(steps::Integer, implicit::ImplicitFineGrid, ops::L2PlusDivAGrad, curr::LevelState{Float64,DeviceMatrix}, k::Int)
    sync_operator!(curr.x.grid, ops)
    check(ccall((:hmg_smoothing_steps, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint), curr.x.grid.ctx, k, steps))
end

"""
    DeviceBaseLevel(grid, A_interior, interior_nodes)

Replaces `BaseLevel(Float64, cholesky(A[interior,interior]), nnodes(base), interior)`
(src/examples/homogenized_coefficients.jl:259-261): the sparse matrix itself is handed over.
"""
struct DeviceBaseLevel
    grid::DeviceGrid
end
function DeviceBaseLevel(g::DeviceGrid, A::SparseMatrixCSC{Float64,Int64}, interior::Vector{Int64})
    GC.@preserve A interior check(ccall((:hmg_set_coarse_matrix, libhmg), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}),
        g.ctx, size(A, 1), A.colptr, A.rowval, A.nzval, interior))
    DeviceBaseLevel(g)
end

# vcycle!(implicit, base, ops, levels, k, steps)   (src/multigrid.jl:73-119): one ccall per V-cycle
function vcycle!(implicit::ImplicitFineGrid, base::DeviceBaseLevel, ops::Vector{<:L2PlusDivAGrad},
                 levels::Vector{LevelState{Float64,DeviceMatrix}}, k::Int, steps = 2)
    sync_operator!(base.grid, ops[k])
    check(ccall((:hmg_vcycle, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}), base.grid.ctx, k, steps, C_NULL))
    nothing
end

# ---- the callers around the V-cycle (src/examples/homogenized_coefficients.jl), on the device ----
# The element subsets of the driver are prefixes 1:n of the magnitude-ordered base mesh (:272).

# rhs_aξ∇v!(b, ∂ϕ∂xᵢs, implicit, σs, ξ)   (:449-474)
function rhs_aξ∇v!(b::DeviceMatrix, ∂ϕ∂xᵢs::Vector{SVector{dim,Float64}}, implicit::ImplicitFineGrid{dim},
                   σs::Vector{SVector{dim,Float64}}, ξ::SVector{dim,Float64}) where {dim}
    xi = collect(ξ)
    GC.@preserve xi check(ccall((:hmg_rhs_axi_grad, libhmg), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), b.grid.ctx, xi, b.which))
    nothing
end

# integrate_first_term(v₀, ∂ϕ∂xᵢs, implicit, subset, ops, σs, ξ)   (:592-632)
function integrate_first_term(v₀::DeviceMatrix, ∂ϕ∂xᵢs::Vector{SVector{dim,Float64}}, implicit::ImplicitFineGrid{dim},
                              subset::Base.OneTo, ops::L2PlusDivAGrad, σs::Vector{SVector{dim,Float64}},
                              ξ::SVector{dim,Float64}) where {dim}
    xi = collect(ξ)
    out = Ref{Float64}(0.0)
    GC.@preserve xi check(ccall((:hmg_integrate_first_term, libhmg), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Ref{Float64}),
                                v₀.grid.ctx, v₀.which, xi, length(subset), out))
    out[]
end

# integrate_terms(vₖ, vₖ₋₁, implicit, subset, ops)   (:634-667)
function integrate_terms(vₖ::DeviceMatrix, vₖ₋₁::DeviceMatrix, implicit::ImplicitFineGrid, subset::Base.OneTo, ops::L2PlusDivAGrad)
    out = Ref{Float64}(0.0)
    check(ccall((:hmg_integrate_terms, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Ref{Float64}),
                vₖ.grid.ctx, vₖ.which, vₖ₋₁.which, length(subset), out))
    out[]
end

# integrate_area(ops, implicit, subset)   (:673-689) -- needs the grid, so it dispatches on a DeviceGrid
function integrate_area(g::DeviceGrid, subset::Base.OneTo)
    out = Ref{Float64}(0.0)
    check(ccall((:hmg_integrate_area, libhmg), Cint, (Ptr{Cvoid}, Int64, Ref{Float64}), g.ctx, length(subset), out))
    out[]
end

# next_rhs!(b, x, implicit, ops)   (:695-713)
function next_rhs!(b::DeviceMatrix, x::DeviceMatrix, implicit::ImplicitFineGrid, ops::L2PlusDivAGrad)
    check(ccall((:hmg_set_lambda, libhmg), Cint, (Ptr{Cvoid}, Float64), b.grid.ctx, ops.λ))
    check(ccall((:hmg_next_rhs, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint), b.grid.ctx, b.which, x.which))
    nothing
end

# SimpleDiffusion(ops, bc, a)   (src/build_local_operators.jl:19-23, product src/apply_local_operators.jl:40-72) is
# L2PlusDivAGrad with sigma = (a, ..., a) and lambda = 0: create the grid with DeviceGrid(implicit, A) below and the
# same entry points serve it.
DeviceGrid(implicit::ImplicitFineGrid{dim}, A::SimpleDiffusion; device::Integer = 0) where {dim} =
    DeviceGrid(implicit, fill(SVector{dim,Float64}(ntuple(_ -> A.a, dim)), nelements(implicit.base)), 0.0; device = device)
function mul!(α::Float64, base::Mesh, A::SimpleDiffusion, x::DeviceMatrix, y::DeviceMatrix)
    check(ccall((:hmg_set_lambda, libhmg), Cint, (Ptr{Cvoid}, Float64), x.grid.ctx, 0.0))
    check(ccall((:hmg_mul, libhmg), Cint, (Ptr{Cvoid}, Cint, Float64, Cint, Cint), x.grid.ctx, x.level, α, x.which, y.which))
    y
end
function local_residual!(implicit::ImplicitFineGrid, A::SimpleDiffusion, curr::LevelState{Float64,DeviceMatrix}, k::Int)
    check(ccall((:hmg_set_lambda, libhmg), Cint, (Ptr{Cvoid}, Float64), curr.x.grid.ctx, 0.0))
    check(ccall((:hmg_local_residual, libhmg), Cint, (Ptr{Cvoid}, Cint), curr.x.grid.ctx, k))
    nothing
end

# local_rhs!(b, implicit)   (src/implicit_fine_grid.jl:391-409): O(Nf * Ne) once per problem -- the reference's own host
# method fills a Matrix, which is uploaded
function local_rhs!(b::DeviceMatrix, implicit::ImplicitFineGrid)
    h = Matrix{Float64}(undef, size(b)...)
    local_rhs!(h, implicit)
    copyto!(b, h)
end

# shrink_level_state(l, nf, n)   (src/examples/homogenized_coefficients.jl:54-60): the new grid lives on an element
# prefix of the old one; l.x[:, OneTo(n)] moves device to device (both contexts on the same GPU).
function shrink_to!(dst::DeviceMatrix, src::DeviceMatrix)
    check(ccall((:hmg_copy_columns_from, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cint),
                dst.grid.ctx, dst.level, dst.which, src.grid.ctx, src.which))
    dst
end

# x[1 : nnodes(refined_mesh(implicit, level)), :]   (export_unknown, :81-87): only the row prefix crosses PCIe
function first_rows(x::DeviceMatrix, nrows::Integer)
    out = Matrix{Float64}(undef, nrows, size(x, 2))
    GC.@preserve out check(ccall((:hmg_download_rows, libhmg), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Float64}, Int64),
                                 x.grid.ctx, x.level, x.which, nrows, out, max(nrows, 1)))
    out
end
function export_unknown(base::Mesh{dim}, implicit::ImplicitFineGrid, x::DeviceMatrix, k::Int, level::Int) where {dim}
    vtk_grid("ahom_$k", construct_full_grid(implicit, level)) do vtk
        vtk_point_data(vtk, first_rows(x, nnodes(refined_mesh(implicit, level)))[:], "v")
    end
end

"""
    PartitionedDeviceGrid(implicit, σs, λ, owner_rank, rank, nranks, nccl_id; device)

One process per GPU (e.g. under MPI.jl): `owner_rank[e]` assigns every coarse element to a rank, `nccl_id` is the
128-byte id produced by `nccl_unique_id()` on rank 0 and broadcast by the launcher.  The state matrices of the
returned grid hold this rank's columns only (`local_elements(grid)`); only interface partial sums and scalars move.
"""
function PartitionedDeviceGrid(implicit::ImplicitFineGrid{dim}, σs::Vector{SVector{dim,Float64}}, λ::Float64,
                               owner_rank::Vector{Int32}, rank::Integer, nranks::Integer, nccl_id::Vector{UInt8};
                               device::Integer = rank) where {dim}
    base = implicit.base
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    nodes = reinterpret(Float64, base.nodes); elems = reinterpret(Int64, base.elements); sig = reinterpret(Float64, σs)
    GC.@preserve nodes elems sig owner_rank nccl_id begin
        check(ccall((:hmg_create_partitioned, libhmg), Cint,
            (Cint, Cint, Int64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Float64, Cint, Cint, Cint, Ptr{Int32}, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
            dim, nlevels(implicit), nelements(base), nnodes(base), nodes, elems, sig, λ, device, rank, nranks, owner_rank, nccl_id, ctx))
    end
    g = DeviceGrid(ctx[], implicit)
    finalizer(x -> (x.ctx == C_NULL || ccall((:hmg_destroy, libhmg), Cint, (Ptr{Cvoid},), x.ctx); x.ctx = C_NULL), g)
    g
end
function nccl_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:hmg_nccl_unique_id, libhmg), Cint, (Ptr{UInt8},), id))
    id
end
function local_elements(g::DeviceGrid)
    out = Vector{Int64}(undef, ccall((:hmg_ne_local, libhmg), Int64, (Ptr{Cvoid},), g.ctx))
    check(ccall((:hmg_local_elements, libhmg), Cint, (Ptr{Cvoid}, Ptr{Int64}), g.ctx, out))
    out
end

end # module
