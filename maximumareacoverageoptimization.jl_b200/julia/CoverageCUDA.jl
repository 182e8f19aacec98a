# CoverageCUDA.jl -- Julia `ccall` binding of libcoverage_cuda (include/coverage_cuda.h).
#
# Drop-in for the reference's objective factory: `CoverageCUDA.createObjective(cells, N, r_max)`
# returns a closure `AreaMaxObjective(x::Vector{Float64})::Float64` with the value semantics of
# src/TDM_STATIC_opt.jl:82-100, so `SetObjective(p, obj)` (src/TDM_STATIC_opt.jl:125) and
# src/FullSimulation.jl:84-97 keep working with only the factory swapped.  A batched entry
# (`objective_batch`) serves a poll-set-at-a-time MADS driver.
#
# NOT EXECUTED in this repository's CI: Julia is not installed in the build container or on the GPU
# box.  The same C entry points are exercised from Python ctypes in tests/test_gpu_parity.py.
module CoverageCUDA

const LIB = get(ENV, "LIBCOVERAGE_CUDA", joinpath(@__DIR__, "..", "libcoverage_cuda.so"))

struct CovError <: Exception
    code::Cint
    msg::String
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    N::Int
    function Handle(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:cov_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, out)
        rc == 0 || throw(CovError(rc, unsafe_string(ccall((:cov_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
        h = new(out[], 0)
        finalizer(h -> (h.ptr == C_NULL || ccall((:cov_destroy, LIB), Cvoid, (Ptr{Cvoid},), h.ptr); h.ptr = C_NULL), h)
        return h
    end
end

check(h::Handle, rc::Cint) =
    rc == 0 || throw(CovError(rc, unsafe_string(ccall((:cov_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))))

"Upload the reference's own point list (`Vector{Vector{Float64}}` of [x, y, area, weight, covered])."
function set_points!(h::Handle, points::Vector{Vector{Float64}}; nx = 100, ny = 100, dx = 5.0, dy = 5.0)
    flat = Vector{Float64}(undef, 5 * length(points))
    @inbounds for (p, pt) in enumerate(points), k in 1:5
        flat[5 * (p - 1) + k] = pt[k]
    end
    GC.@preserve flat check(h, ccall((:cov_set_points, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Float64, Float64),
        h.ptr, flat, length(points), nx, ny, dx, dy))
end

"createPOI(dx, dy, nx, ny) built on the device (src/AreaCoverageCalculation.jl:11-21)."
set_grid_full!(h::Handle, nx, ny, dx, dy) =
    check(h, ccall((:cov_set_grid_full, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Float64, Float64), h.ptr, nx, ny, dx, dy))

"rmvCoveredPOI on the device-resident store (src/CellFunctions.jl:81-108); returns entries removed."
function remove_covered!(h::Handle, xyR::Vector{Float64})
    removed = Ref{Int64}(0)
    GC.@preserve xyR check(h, ccall((:cov_remove_covered, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Int64}),
        h.ptr, xyR, length(xyR) ÷ 3, removed))
    return removed[]
end

"update_POI on the device-resident store (src/CellFunctions.jl:59-77): append list entries, duplicates add up."
function add_points!(h::Handle, points::Vector{Vector{Float64}})
    flat = Vector{Float64}(undef, 5 * length(points))
    @inbounds for (p, pt) in enumerate(points), k in 1:5
        flat[5 * (p - 1) + k] = pt[k]
    end
    GC.@preserve flat check(h, ccall((:cov_add_points, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64),
        h.ptr, flat, length(points)))
end

"""
    fire_init!(h, state; dx = 5.0, dy = 5.0, push_initial = true)
    fire_step!(h, wind_speed, wind_direction, prob_spread; seed = 0, step = 0, append = true)

The forest-fire automaton of src/DynamicArea.jl:17-86 on the device: `state` is the `nx x ny` `UInt8` grid
(0 EMPTY, 1 TREE, 2 FIRE); every step pushes the newly ignited cells straight into the cell store that the
objective reads (no xlsx hand-off). `fire_step!` returns the number of entries pushed.
"""
function fire_init!(h::Handle, state::Matrix{UInt8}; dx = 5.0, dy = 5.0, push_initial = true)
    nx, ny = size(state)
    GC.@preserve state check(h, ccall((:cov_fire_init, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Float64, Float64, Ptr{UInt8}, Int32), h.ptr, nx, ny, dx, dy, state, push_initial ? 1 : 0))
end
function fire_step!(h::Handle, wind_speed, wind_direction, prob_spread; seed = 0, step = 0, append = true)
    pushed = Ref{Int64}(0)
    check(h, ccall((:cov_fire_step, LIB), Cint,
        (Ptr{Cvoid}, Float64, Float64, Float64, UInt64, Int64, Int32, Ref{Int64}),
        h.ptr, wind_speed, wind_direction, prob_spread, seed, step, append ? 1 : 0, pushed))
    return pushed[]
end

"Exact area of the union of the discs of every column of `X` (3N x B, [x; y; R]); no grid involved."
function union_area(h::Handle, X::Matrix{Float64})
    N = size(X, 1) ÷ 3; B = size(X, 2)
    area = Vector{Float64}(undef, B)
    GC.@preserve X area check(h, ccall((:cov_union_area_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}), h.ptr, X, B, N, area))
    return area
end

"Captured variables of createObjective / create_cons3 / cons8 / cons7."
function set_params!(h::Handle, N::Integer, r_max::Vector{Float64}; penalty = 1e5,
                     prev::Union{Nothing,Vector{Float64}} = nothing, d_lim::Vector{Float64} = fill(10.0, N),
                     FOV = 100 / 180 * π, sep_min = 0.0, use_cons7 = false)
    prevp = prev === nothing ? Ptr{Float64}(C_NULL) : pointer(prev)
    GC.@preserve r_max prev d_lim check(h, ccall((:cov_set_params, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Int32),
        h.ptr, N, r_max, penalty, prevp, d_lim, tan(FOV / 2), sep_min, use_cons7 ? 1 : 0))
    h.N = N
end

"One candidate: AreaMaxObjective(x) (src/TDM_STATIC_opt.jl:83-98)."
function eval_one(h::Handle, x::Vector{Float64})
    obj = Ref{Float64}(0.0)
    GC.@preserve x check(h, ccall((:cov_eval_one, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), h.ptr, x, obj))
    return obj[]
end

"A whole poll set: X is 3N x B (one candidate [x;y;R] per COLUMN, i.e. candidate-major in memory).
Allocates plain (pageable) outputs; for large or repeated batches use `BatchBuffers` + `objective_batch!` (pinned)."
function objective_batch(h::Handle, X::Matrix{Float64})
    B = size(X, 2)
    obj = Vector{Float64}(undef, B); count = Vector{Int64}(undef, B); feasible = Vector{UInt8}(undef, B)
    GC.@preserve X obj count feasible check(h, ccall((:cov_eval_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{UInt8}),
        h.ptr, X, B, obj, count, feasible))
    return obj, count, feasible
end

"""
    pinned_matrix(h, rows, cols) / pinned_vector(h, T, n)

Page-locked host arrays owned by the library (`cov_host_alloc`), wrapped without a copy. `cov_eval_batch` DMAs pinned
buffers in place; a plain `Matrix` goes through the library's pinned staging instead (measured on B200, 1 M candidates
x 5 UAVs: 2.45 ms pinned vs 3.3 ms pageable per call). The arrays stay valid until `free_pinned!` or until the handle
is destroyed -- keep the handle alive as long as they are in use.
"""
function pinned_vector(h::Handle, ::Type{T}, n::Integer) where {T}
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(h, ccall((:cov_host_alloc, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), h.ptr, max(n * sizeof(T), 1), out))
    return unsafe_wrap(Array, Ptr{T}(out[]), (Int(n),); own = false)
end
function pinned_matrix(h::Handle, rows::Integer, cols::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(h, ccall((:cov_host_alloc, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), h.ptr, max(rows * cols * 8, 1), out))
    return unsafe_wrap(Array, Ptr{Float64}(out[]), (Int(rows), Int(cols)); own = false)
end
free_pinned!(h::Handle, a::Array) = check(h, ccall((:cov_host_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, pointer(a)))

"Reusable pinned buffers for poll sets of up to `B` candidates: fill `X[:, 1:b]`, call `objective_batch!(h, bufs, b)`."
struct BatchBuffers
    X::Matrix{Float64}        # 3N x B, one candidate per column
    obj::Vector{Float64}
    count::Vector{Int64}
    feasible::Vector{UInt8}
end
BatchBuffers(h::Handle, B::Integer) = BatchBuffers(pinned_matrix(h, 3 * h.N, B), pinned_vector(h, Float64, B),
                                                   pinned_vector(h, Int64, B), pinned_vector(h, UInt8, B))
function objective_batch!(h::Handle, b::BatchBuffers, n::Integer = size(b.X, 2))
    check(h, ccall((:cov_eval_batch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{UInt8}),
        h.ptr, b.X, n, b.obj, b.count, b.feasible))
    return view(b.obj, 1:n), view(b.count, 1:n), view(b.feasible, 1:n)
end

"""
    objective_batch_best!(h, bufs, n; barrier = true) -> (obj, count, feasible, best_obj, best_column)

`objective_batch!` plus the poll winner reduced on the device (`cov_eval_batch_best`): the 16 bytes a process
contributes when the candidates of one poll are sharded over several GPUs (one Julia process per GPU, e.g. under
MPI.jl or Distributed: `MPI.Allgather((best_obj, first_column + best_column))`, smallest objective wins, ties to the
smallest column). `best_column` is 1-based, 0 when no candidate is feasible.
"""
function objective_batch_best!(h::Handle, b::BatchBuffers, n::Integer = size(b.X, 2); barrier = true)
    bo = Ref{Float64}(Inf); bi = Ref{Int64}(-1)
    check(h, ccall((:cov_eval_batch_best, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{UInt8}, Int32, Ref{Float64}, Ref{Int64}),
        h.ptr, b.X, n, b.obj, b.count, b.feasible, barrier ? 1 : 0, bo, bi))
    return view(b.obj, 1:n), view(b.count, 1:n), view(b.feasible, 1:n), bo[], bi[] + 1
end

_pack_code(::Type{Float32}) = Int32(1)
_pack_code(::Type{Int32}) = Int32(2)
_pack_code(::Type{Int16}) = Int32(3)

"Pinned `rows x cols` matrix of packed candidate values (`Int16`, `Int32` or `Float32`) for `objective_batch_packed`."
function pinned_packed(h::Handle, ::Type{T}, rows::Integer, cols::Integer) where {T<:Union{Int16,Int32,Float32}}
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(h, ccall((:cov_host_alloc, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), h.ptr, max(rows * cols * sizeof(T), 1), out))
    return unsafe_wrap(Array, Ptr{T}(out[]), (Int(rows), Int(cols)); own = false)
end

"""
    objective_batch_packed(h, Q, granularity = 1.0; winner = false, barrier = true)

Poll / search sets whose trial points sit on the MADS mesh, sent PACKED (`cov_eval_batch_packed`): `Q` is `3N x B` of
`Int16` or `Int32` mesh indices (candidate = `q * granularity`; granularity 1.0 on every variable in
`TDM_STATIC_opt.optimize`, src/TDM_STATIC_opt.jl:131-137) or of `Float32` values. A quarter or half of the bytes cross
PCIe, the kernels see the same `Float64` candidates, the results are bit for bit those of `objective_batch` on
`Float64.(Q) .* granularity` (measured on B200, 1 M candidates x 5 UAVs from pinned buffers: 2.40 ms -> 1.11 ms per
call with `Int16`). `Q` from `pinned_packed(h, T, 3N, B)` is DMA'd in place. With `winner = true` also returns
`(best_obj, best_column)` like `objective_batch_best!`.
"""
function objective_batch_packed(h::Handle, Q::Matrix{T}, granularity::Real = 1.0; winner = false, barrier = true) where {T<:Union{Int16,Int32,Float32}}
    B = size(Q, 2)
    obj = Vector{Float64}(undef, B); count = Vector{Int64}(undef, B); feasible = Vector{UInt8}(undef, B)
    if !winner
        GC.@preserve Q obj count feasible check(h, ccall((:cov_eval_batch_packed, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{UInt8}, Int32, Ptr{Float64}, Ptr{Int64}),
            h.ptr, Q, _pack_code(T), Float64(granularity), B, obj, count, feasible, 0, C_NULL, C_NULL))
        return obj, count, feasible
    end
    bo = Ref{Float64}(Inf); bi = Ref{Int64}(-1)
    GC.@preserve Q obj count feasible check(h, ccall((:cov_eval_batch_packed, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{UInt8}, Int32, Ref{Float64}, Ref{Int64}),
        h.ptr, Q, _pack_code(T), Float64(granularity), B, obj, count, feasible, barrier ? 1 : 0, bo, bi))
    return obj, count, feasible, bo[], bi[] + 1
end

"The store's weight classes in the device's numbering (the `class_count` columns of `cov_eval_batch_ex`)."
function class_weights(h::Handle)
    w = zeros(Float64, 4)
    check(h, ccall((:cov_get_class_weights, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), h.ptr, w, 4))
    return w
end

"""
    create_cons1_progressive(N, r_max) / create_cons2_progressive / create_cons3_progressive

The progressive constraints of src/TDM_Constraints.jl:182-221 (`x -> Real`, for `AddProgressiveConstraint`):
`sum_i max(R_i - r_max_i, 0.0)`, and the same term for UAV 2 / UAV 3 alone. `COV_OPT_PROGRESSIVE_INDEX` (10) selects
which one the `progressive` output of `cov_eval_batch_ex` carries; `.batch(X)`-style use: `progressive_batch`.
"""
function progressive_batch(h::Handle, X::Matrix{Float64}, which::Integer)
    B = size(X, 2)
    obj = Vector{Float64}(undef, B); prog = Vector{Float64}(undef, B)
    check(h, ccall((:cov_set_option, LIB), Cint, (Ptr{Cvoid}, Cint, Int64), h.ptr, 10, which))
    GC.@preserve X obj prog check(h, ccall((:cov_eval_batch_ex, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}),
        h.ptr, X, B, obj, C_NULL, C_NULL, C_NULL, prog))
    return prog
end
function _create_progressive(which::Integer, N::Integer, r_max; handle::Handle = Handle())
    set_grid_full!(handle, 1, 1, 1.0, 1.0)          # constraints do not look at the cell store
    set_params!(handle, N, collect(Float64, r_max); penalty = 0.0)
    return x -> progressive_batch(handle, reshape(collect(Float64, x), :, 1), which)[1]
end
create_cons1_progressive(N, r_max; kw...) = _create_progressive(0, N, r_max; kw...)
create_cons2_progressive(N, r_max; kw...) = _create_progressive(2, N, r_max; kw...)
create_cons3_progressive(N, r_max; kw...) = _create_progressive(3, N, r_max; kw...)

"Poll winner with the extreme barrier: (best objective, 1-based column) or (Inf, 0)."
function argmin_batch(h::Handle, X::Matrix{Float64}; barrier = true)
    bo = Ref{Float64}(Inf); bi = Ref{Int64}(-1)
    GC.@preserve X check(h, ccall((:cov_argmin, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ref{Float64}, Ref{Int64}), h.ptr, X, size(X, 2), barrier ? 1 : 0, bo, bi))
    return bo[], bi[] + 1
end

"""
    mads_solve(h, x0; N_iter = 100, granularity = 1.0, seed = 0)

The whole MADS solve inside the library (same settings as `TDM_STATIC_opt.optimize`, src/TDM_STATIC_opt.jl:118-222;
the constraints given to `set_params!` act as extreme barrier). Returns `(result, objective, runtime_seconds)`.
"""
function mads_solve(h::Handle, x0::Vector{Float64}; N_iter = 100, granularity = 1.0, seed = 0)
    out = similar(x0); obj = Ref{Float64}(0.0); stats = zeros(Int64, 4)
    t = @elapsed GC.@preserve x0 out stats check(h, ccall((:cov_mads_solve, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, UInt64, Ptr{Float64}, Ref{Float64}, Ptr{Int64}),
        h.ptr, x0, N_iter, granularity, seed, out, obj, stats))
    return out, obj[], t
end

"""
    createObjective(cells, N, r_max; handle = Handle())

Same signature and value as the reference's `TDM_STATIC_opt.createObjective` (src/TDM_STATIC_opt.jl:82):
uploads `cells.points_of_interest` once and returns `AreaMaxObjective(x)`.
"""
function createObjective(cells, N, r_max; handle::Handle = Handle(), nx = 100, ny = 100, dx = 5.0, dy = 5.0)
    set_points!(handle, cells.points_of_interest; nx = nx, ny = ny, dx = dx, dy = dy)
    set_params!(handle, N, collect(Float64, r_max))
    AreaMaxObjective(x) = eval_one(handle, x)
    return AreaMaxObjective
end

end # module
