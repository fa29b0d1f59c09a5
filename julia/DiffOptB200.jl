# DiffOptB200.jl -- Julia host side of the B200 sensitivity hot path (thin `ccall` layer over
# include/diffopt_b200.h).  NOT executed in the build image (no Julia there); it is the binding a
# DiffOpt.jl maintainer adds, kept in sync with the C header and exercised through the identical C ABI
# by the Python ctypes mirror in diffopt.jl_b200/ (see INTEGRATION.md).
#
# Plug points used (reference tree andrewrosemberg/DiffOpt.jl v0.5.0):
#   * DiffOpt.ModelConstructor            src/moi_wrapper.jl:504-514, consumed by _diff :619-657
#   * QuadraticProgram.LinearAlgebraSolver + solve_system(solver, LHS, RHS, iterative)
#                                         src/QuadraticProgram/QuadraticProgram.jl:476-502
module DiffOptB200

import DiffOpt
import LinearAlgebra
import MathOptInterface as MOI
import SparseArrays

const LIB = get(ENV, "DIFFOPT_B200_LIB", "libdiffopt_b200")
const HOST = Cint(0)
const DEVICE = Cint(1)

struct B200Error <: Exception
    code::Int32
    msg::String
end

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:diffopt_b200_create, LIB), Int32, (Int32, Ptr{Ptr{Cvoid}}), device, ref)
        rc == 0 || throw(B200Error(rc, "diffopt_b200_create failed (no B200 / no CUDA device: there is no CPU fallback)"))
        ctx = new(ref[])
        finalizer(c -> ccall((:diffopt_b200_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

last_error(ctx::Context) = unsafe_string(ccall((:diffopt_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle))

# rc < 0: bad argument / CUDA error -> error(...) like the reference's generic failures;
# rc > 0: LAPACK-style info -> SingularException, what `LHS \ RHS` throws in the reference.
function check(ctx::Context, rc::Int32)
    rc < 0 && throw(B200Error(rc, last_error(ctx)))
    rc > 0 && throw(LinearAlgebra.SingularException(Int(rc)))
    return
end

# ------------------------------------------------------------------------------------------------
# (1) Narrow plug point: a linear solver for QuadraticProgram.Model
#     model.linear_solver = B200Solver(ctx);  solve_system is called from
#     reverse_differentiate! (:335, LHS) and forward_differentiate! (:438, LHS')
# ------------------------------------------------------------------------------------------------
struct B200Solver
    ctx::Context
    atol::Float64
    btol::Float64
    conlim::Float64
end
B200Solver(ctx::Context) = B200Solver(ctx, sqrt(eps()), sqrt(eps()), 1 / sqrt(eps()))  # IterativeSolvers.lsqr defaults

_csc(A::SparseArrays.SparseMatrixCSC{Float64,Int}) = (A, Cint(0))
_csc(A::LinearAlgebra.Adjoint{Float64,<:SparseArrays.SparseMatrixCSC{Float64,Int}}) = (parent(A), Cint(1))

function DiffOpt.QuadraticProgram.solve_system(s::B200Solver, LHS, RHS::AbstractVector, iterative::Bool)
    A, trans = _csc(LHS)
    rhs = Vector{Float64}(RHS)           # forward mode may hand over a SparseVector (:493-496)
    n = size(A, 1)
    x = Vector{Float64}(undef, n)
    GC.@preserve A rhs x begin
        rc = if iterative
            # LP branch: IterativeSolvers.lsqr(LHS, RHS) -> persistent device LSQR on the CSC matrix
            ccall((:diffopt_b200_lsqr_csc, LIB), Int32,
                  (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Ptr{Float64},
                   Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Int32),
                  s.ctx.handle, size(A, 1), size(A, 2), A.colptr, A.rowval, A.nzval, trans, rhs,
                  s.atol, s.btol, s.conlim, max(size(A)...), x, C_NULL, HOST)
        else
            # `LHS \ RHS`: pivoted LU on the device
            ccall((:diffopt_b200_kkt_solve_csc, LIB), Int32,
                  (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Int32),
                  s.ctx.handle, n, A.colptr, A.rowval, A.nzval, trans, 1, rhs, x, HOST)
        end
    end
    check(s.ctx, rc)
    return x
end

"""
    SparseFactorization(ctx, LHS)            # LHS::SparseMatrixCSC or its Adjoint

One device factorisation of a large sparse KKT matrix (RCM ordering + banded LU with partial pivoting), reused for any
number of right-hand sides: `F \\ RHS` with `RHS::Matrix` (N x nrhs).  Replaces the per-direction `LHS' \\ RHS` of
`forward_differentiate!` (QuadraticProgram.jl:438) when many directions are differentiated against one solution.
"""
struct SparseFactorization
    ctx::Context
    n::Int
    bandwidth::Int
    function SparseFactorization(ctx::Context, LHS)
        A, trans = _csc(LHS)
        bw = Ref{Int64}(0)
        GC.@preserve A begin
            rc = ccall((:diffopt_b200_sparse_setup, LIB), Int32,
                       (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Ptr{Int64}),
                       ctx.handle, size(A, 1), A.colptr, A.rowval, A.nzval, trans, bw)
        end
        check(ctx, rc)
        return new(ctx, size(A, 1), Int(bw[]))
    end
end

function Base.:\(F::SparseFactorization, RHS::StridedVecOrMat{Float64})
    X = similar(RHS)
    GC.@preserve RHS X begin
        rc = ccall((:diffopt_b200_sparse_solve, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int32),
                   F.ctx.handle, size(RHS, 2), RHS, X, HOST)
    end
    check(F.ctx, rc)
    return X
end

"ModelConstructor that keeps the B200 solver attached although the wrapper rebuilds the backend (moi_wrapper.jl:619-657)."
function qp_model_constructor(ctx::Context)
    return () -> begin
        m = DiffOpt.QuadraticProgram.Model()
        m.linear_solver = B200Solver(ctx)
        m
    end
end
# usage:  MOI.set(model, DiffOpt.ModelConstructor(), DiffOptB200.qp_model_constructor(ctx))

# ------------------------------------------------------------------------------------------------
# (2) Batched QP sensitivities (OptNet-style layers): B independent instances, dense column-major
#     arrays exactly as Julia stores them: Q[n,n,B], G[m,n,B], A[p,n,B], h[m,B], z[n,B], lam[m,B], nu[p,B]
#     (lam, nu are the reference's stored duals, i.e. NEGATED MOI duals, QuadraticProgram.jl:156-180).
#     One call = create_LHS_matrix (:256-282) + forward (:357-446) + reverse (:316-351) for every instance.
# ------------------------------------------------------------------------------------------------
_p(a::Nothing) = Ptr{Float64}(C_NULL)
_p(a::Array{Float64}) = pointer(a)

function qp_batch_solve(ctx::Context, Q, G, A, h, z, lam, nu;
                        dQ = nothing, dq = nothing, dG = nothing, dh = nothing, dA = nothing, db = nothing,
                        dl_dz = nothing)
    n, B = size(z)
    m = size(lam, 1)
    p = size(nu, 1)
    N = n + m + p
    fwd = any(!isnothing, (dQ, dq, dG, dh, dA, db)) ? Matrix{Float64}(undef, N, B) : nothing
    rev = dl_dz === nothing ? nothing : Matrix{Float64}(undef, N, B)
    info = zeros(Int32, B)
    GC.@preserve Q G A h z lam nu dQ dq dG dh dA db dl_dz fwd rev info begin
        rc = ccall((:diffopt_b200_qp_batch_solve, LIB), Int32,
                   (Ptr{Cvoid}, Int64, Int32, Int32, Int32,
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32),
                   ctx.handle, B, n, m, p, _p(Q), _p(G), _p(A), _p(h), _p(z), _p(lam), _p(nu),
                   _p(dQ), _p(dq), _p(dG), _p(dh), _p(dA), _p(db), _p(dl_dz), _p(fwd), _p(rev), info, HOST)
    end
    check(ctx, rc)
    split(x) = x === nothing ? nothing : (dz = x[1:n, :], dλ = x[n+1:n+m, :], dν = x[n+m+1:end, :])
    return (forward = split(fwd), reverse = split(rev), info = info)
end

# Stream-ordered form for device-resident batches (pointers of CUDA.jl arrays, `pointer(x)` converted to Ptr{Float64}):
# enqueue any number of batches, then `synchronize(ctx)` waits and throws for the status of the last one.
function qp_batch_solve_async(ctx::Context, B::Integer, n::Integer, m::Integer, p::Integer,
                              Q::Ptr{Float64}, G::Ptr{Float64}, A::Ptr{Float64}, h::Ptr{Float64}, z::Ptr{Float64},
                              lam::Ptr{Float64}, nu::Ptr{Float64}, dQ::Ptr{Float64}, dq::Ptr{Float64},
                              dG::Ptr{Float64}, dh::Ptr{Float64}, dA::Ptr{Float64}, db::Ptr{Float64},
                              dl_dz::Ptr{Float64}, fwd::Ptr{Float64}, rev::Ptr{Float64}, info::Ptr{Int32})
    rc = ccall((:diffopt_b200_qp_batch_solve_async, LIB), Int32,
               (Ptr{Cvoid}, Int64, Int32, Int32, Int32,
                Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
               ctx.handle, B, n, m, p, Q, G, A, h, z, lam, nu, dQ, dq, dG, dh, dA, db, dl_dz, fwd, rev, info)
    check(ctx, rc)
    return nothing
end

synchronize(ctx::Context) = check(ctx, ccall((:diffopt_b200_synchronize, LIB), Int32, (Ptr{Cvoid},), ctx.handle))

# ------------------------------------------------------------------------------------------------
# (3) ConicProgram backend: _gradient_cache / forward / reverse (src/ConicProgram/ConicProgram.jl:172-394)
#     cone_type: 0 Zeros, 1 Nonnegatives, 2 SecondOrderCone, 3 PositiveSemidefiniteConeTriangle (row order of A)
# ------------------------------------------------------------------------------------------------
function conic_setup(ctx::Context, A::SparseArrays.SparseMatrixCSC{Float64,Int}, b, c, x, s, y,
                     cone_type::Vector{Int32}, cone_dim::Vector{Int64})
    m, n = size(A)
    GC.@preserve A b c x s y cone_type cone_dim begin
        rc = ccall((:diffopt_b200_conic_setup, LIB), Int32,
                   (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int64}, Int32),
                   ctx.handle, n, m, A.colptr, A.rowval, A.nzval, b, c, x, s, y, length(cone_type), cone_type, cone_dim, HOST)
    end
    check(ctx, rc)
end

"reverse_differentiate! (:336-394): g = lsqr(M, [dx; 0; -x'dx]); returns (g, dc, db) with the getters of :396-428."
function conic_reverse(ctx::Context, n::Int, m::Int, dx::Vector{Float64};
                       atol = sqrt(eps()), btol = sqrt(eps()), conlim = 1 / sqrt(eps()), maxiter = n + m + 1)
    g = Vector{Float64}(undef, n + m + 1); dc = Vector{Float64}(undef, n); db = Vector{Float64}(undef, m)
    stats = zeros(4)
    GC.@preserve dx g dc db stats begin
        rc = ccall((:diffopt_b200_conic_reverse, LIB), Int32,
                   (Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32),
                   ctx.handle, dx, atol, btol, conlim, maxiter, g, dc, db, stats, HOST)
    end
    check(ctx, rc)
    return (g = g, dc = dc, db = db, istop = Int(stats[1]), iterations = Int(stats[2]))
end

"forward_differentiate! (:257-334): dA as COO triplets exactly as packed by the reference (un-negated, :296-305)."
function conic_forward(ctx::Context, n::Int, m::Int, dA_rows::Vector{Int64}, dA_cols::Vector{Int64}, dA_vals::Vector{Float64},
                       db, dc; atol = sqrt(eps()), btol = sqrt(eps()), conlim = 1 / sqrt(eps()), maxiter = n + m + 1)
    dx = Vector{Float64}(undef, n); dz = Vector{Float64}(undef, n + m + 1); stats = zeros(4)
    GC.@preserve dA_rows dA_cols dA_vals db dc dx dz stats begin
        rc = ccall((:diffopt_b200_conic_forward, LIB), Int32,
                   (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32),
                   ctx.handle, length(dA_vals), dA_rows, dA_cols, dA_vals, _p(db), _p(dc), atol, btol, conlim, maxiter, dx, dz, stats, HOST)
    end
    check(ctx, rc)
    return (dx = dx, dz = dz, iterations = Int(stats[2]))
end

end # module
