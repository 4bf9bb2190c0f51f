# LagrangianVoronoiB200.jl -- the reference-side binding of liblvb200.so.
#
# This is the shim a maintainer of LagrangianVoronoi.jl adds to run the per-timestep
# mesh-and-pressure hot path on a B200: it keeps every public signature of the package
# (VoronoiGrid, remesh!, PressureSolver, find_pressure!, the examples, the VTK/pvd IO) and
# replaces only the bodies of
#     remesh!(grid)                                   src/voronoigrid.jl:89-108
#     find_pressure!(solver, dt, niter; boundary_velocity)   src/pressure.jl:215-225
#     mul!(y, A::PressureOperator, x)                 src/pressure.jl:119-130
# by `ccall`s into the C ABI declared in include/lv_capi.h.
#
# NOTE: Julia is not installed in the authoring container, so this file has never been
# executed; it is written against include/lv_capi.h and mirrors, call for call, the Python
# host in lagrangianvoronoi.jl_b200/host.py, which IS exercised by the test-suite.
#
# Usage (examples stay unmodified):
#     include("src/LagrangianVoronoi.jl"); using .LagrangianVoronoi
#     include("julia/LagrangianVoronoiB200.jl"); LagrangianVoronoiB200.enable!(device = 0)
#     include("examples/gresho.jl"); gresho.main()
module LagrangianVoronoiB200

using ..LagrangianVoronoi
using ..LagrangianVoronoi: VoronoiGrid, VoronoiPolygon, PressureSolver, PressureOperator, ThreadedVec,
                           Edge, RealVector, Rectangle, FastVector
import LinearAlgebra: mul!

const LIB = get(ENV, "LVB200_LIB", "liblvb200.so")

# ---- include/lv_capi.h ---------------------------------------------------------------------
struct LvGridDesc            # typedef struct LvGridDesc
    dr::Cdouble; h::Cdouble; r_max::Cdouble
    xperiodic::Int32; yperiodic::Int32
    bmin::NTuple{2,Cdouble}; bmax::NTuple{2,Cdouble}
end
const LV_OK, LV_EINVAL, LV_EDESTROYED, LV_ENAN, LV_ECUDA, LV_ECAPACITY = Int32.(0:5)
const LV_SOLVER_CG, LV_SOLVER_MINRES, LV_SOLVER_PCG = Int32(0), Int32(1), Int32(2)

# `Edge` (src/geometry.jl:82-87) is isbits {SVector{2,Float64}, SVector{2,Float64}, Int64} = 40 bytes,
# exactly `LvEdge`; a Vector{Edge} can be handed to the library as LvEdge*.
@assert sizeof(Edge) == 40

mutable struct DeviceContext
    handle::Ptr{Cvoid}
    xy::Vector{RealVector}       # gathered positions (pinned when CUDA.jl is around; plain otherwise)
    rowptr::Vector{Int64}        # n+1 offsets into `edges`
    edges::Vector{Edge}          # flat view of every p.edges, filled by one device->host copy
    area::Vector{Float64}
    centroid::Vector{RealVector}
    fields::Dict{Symbol,Vector}  # gather / scatter staging for the pressure solve
    pending::Bool                # pipelined mode: the last remesh! is still in flight (see sync_mesh!)
end

# Page-lock a staging vector so that the library can DMA from / store into it directly (full PCIe rate, and the direct
# stores described at lv_set_async_edges in include/lv_capi.h).  Re-registered when `resize!` moved the buffer.
const CUDART = get(ENV, "LVB200_CUDART", "libcudart.so")
const PINNED = IdDict{Any,Tuple{Ptr{Cvoid},Int}}()
function pin!(v::Vector)
    p, nb = Ptr{Cvoid}(pointer(v)), sizeof(v)
    old = get(PINNED, v, (C_NULL, 0))
    old == (p, nb) && return v
    old[1] != C_NULL && ccall((:cudaHostUnregister, CUDART), Int32, (Ptr{Cvoid},), old[1])
    nb > 0 && ccall((:cudaHostRegister, CUDART), Int32, (Ptr{Cvoid}, Csize_t, UInt32), p, nb, 0x01) == 0 && (PINNED[v] = (p, nb))
    return v
end

const CONTEXTS = IdDict{Any,DeviceContext}()
const DEVICE = Ref{Int32}(0)
const ENABLED = Ref(false)
const PIPELINED = Ref(false)   # pipelined!(true): deferred remesh + 20 B/edge wire format (lv_set_async_edges(h, 3))

function check(ctx, status::Int32)
    status == LV_OK && return
    msg = unsafe_string(ccall((:lv_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx === nothing ? C_NULL : ctx.handle))
    # same exceptions as the reference
    status == LV_EDESTROYED && throw("The Voronoi Mesh has been destroyed.")      # voronoigrid.jl:64
    status == LV_ENAN && throw("Velocity field invalidated.")                      # move.jl:25
    status == LV_EINVAL && occursin("h must be positive", msg) && throw(ArgumentError("h must be positive"))
    error("liblvb200 status $status: $msg")
end

# one device context per grid, created lazily from the grid's own fields (voronoigrid.jl:14-25)
function context(grid::VoronoiGrid)
    get!(CONTEXTS, grid) do
        b = grid.boundary_rect
        desc = Ref(LvGridDesc(grid.dr, grid.h, sqrt(grid.rr_max), Int32(grid.xperiodic), Int32(grid.yperiodic),
                              (b.xmin[1], b.xmin[2]), (b.xmax[1], b.xmax[2])))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:lv_create, LIB), Int32, (Ref{LvGridDesc}, Int32, Ref{Ptr{Cvoid}}), desc, DEVICE[], h)
        st == LV_OK || check(nothing, st)
        ctx = DeviceContext(h[], RealVector[], Int64[], Edge[], Float64[], RealVector[], Dict{Symbol,Vector}(), false)
        PIPELINED[] && check(ctx, ccall((:lv_set_async_edges, LIB), Int32, (Ptr{Cvoid}, Int32), h[], 3))
        finalizer(c -> ccall((:lv_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), ctx)
        ctx
    end
end

# ---- remesh!(grid)  src/voronoigrid.jl:89-108 --------------------------------------------------
function remesh_b200!(grid::VoronoiGrid)
    ctx = context(grid)
    n = length(grid.polygons)
    resize!(ctx.xy, n); resize!(ctx.rowptr, n + 1); resize!(ctx.area, n); resize!(ctx.centroid, n)
    length(ctx.edges) < 7n + 64 && resize!(ctx.edges, 7n + 64)
    foreach(pin!, (ctx.xy, ctx.rowptr, ctx.area, ctx.centroid, ctx.edges))
    Threads.@threads for i in 1:n
        @inbounds ctx.xy[i] = grid.polygons[i].x
    end
    # examples/piston.jl:43-47 mutates the rectangles between remeshes
    b, c = grid.boundary_rect, grid.cropping_rect
    check(ctx, ccall((:lv_set_rects, LIB), Int32, (Ptr{Cvoid}, Ref{RealVector}, Ref{RealVector}, Ref{RealVector}, Ref{RealVector}),
                     ctx.handle, b.xmin, b.xmax, c.xmin, c.xmax))
    nnz = Ref{Int64}(0)
    st = ccall((:lv_remesh, LIB), Int32,
               (Ptr{Cvoid}, Int64, Ptr{RealVector}, Ptr{Int64}, Ptr{Edge}, Int64, Ref{Int64}, Ptr{Float64}, Ptr{RealVector}),
               ctx.handle, n, ctx.xy, ctx.rowptr, ctx.edges, length(ctx.edges), nnz, ctx.area, ctx.centroid)
    if st == LV_ECAPACITY && nnz[] > length(ctx.edges)
        resize!(ctx.edges, nnz[])
        st = ccall((:lv_mesh_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Edge}, Int64, Ptr{Float64}, Ptr{RealVector}),
                   ctx.handle, ctx.rowptr, ctx.edges, length(ctx.edges), ctx.area, ctx.centroid)
    end
    check(ctx, st)
    if PIPELINED[]          # lv_set_async_edges(h, 3): the clip kernel is only queued; the edge view is scattered by sync_mesh!
        ctx.pending = true
        return
    end
    scatter_edges!(grid, ctx)
    return
end

# Pipelined mode (include/lv_capi.h, lv_set_async_edges(h, 3)): remesh! returns with its kernel queued, the next remesh! /
# find_pressure! uploads while it runs, and the mesh arrives as 20 B/edge expanded by the library's host threads.  The first
# consumer of p.edges (pressure_step!, export_grid, ...) calls sync_mesh!(grid); errors of the deferred remesh surface there
# or at the next device call, with the reference's messages.
function pipelined!(on::Bool = true)
    for ctx in values(CONTEXTS)
        check(ctx, ccall((:lv_set_async_edges, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.handle, on ? 3 : 0))
    end
    PIPELINED[] = on
end
function sync_mesh!(grid::VoronoiGrid)
    ctx = context(grid)
    ctx.pending || return
    check(ctx, ccall((:lv_mesh_wait, LIB), Int32, (Ptr{Cvoid},), ctx.handle))
    ctx.pending = false
    scatter_edges!(grid, ctx)
end

function scatter_edges!(grid::VoronoiGrid, ctx::DeviceContext)
    n = length(grid.polygons)
    # p.edges keeps its concrete type FastVector{Edge} (polygon.jl:23-27): copy each row into the
    # polygon's own buffer.  (Zero-copy alternative, a two-line change in celldefs.jl:24 and
    # polygon.jl:33: make `edges` a `SubArray` of ctx.edges -- see INTEGRATION.md.)
    Threads.@threads for i in 1:n
        @inbounds begin
            p = grid.polygons[i]
            lo, hi = ctx.rowptr[i] + 1, ctx.rowptr[i+1]
            k = hi - lo + 1
            length(p.edges.data) < k && resize!(p.edges.data, k)
            copyto!(p.edges.data, 1, ctx.edges, lo, k)
            p.edges.last = k
        end
    end
    return
end

# ---- find_pressure!  src/pressure.jl:215-225 ---------------------------------------------------
# boundary_velocity(midpoint(e), e.label) is a Julia closure (pressure.jl:182); no callback crosses the C ABI, so it
# is evaluated HERE for every boundary edge, in the order boundaries(p) yields them polygon by polygon (= the numbering
# of lv_boundary_edges: the edge view is already on the host after remesh!).  When the values are constant along every
# wall (all the reference's examples, e.g. examples/piston.jl:122-127) the four per-wall constants go down and the per-edge
# array stays empty; otherwise the per-edge values do.
function wall_velocities(grid::VoronoiGrid, boundary_velocity)
    wall = fill(VEC0, 4)
    seen = falses(4)
    constant = true
    edge = RealVector[]
    for p in grid.polygons, e in LagrangianVoronoi.boundaries(p)
        vbc = boundary_velocity(LagrangianVoronoi.midpoint(e), e.label)
        push!(edge, vbc)
        k = -e.label
        if 1 <= k <= 4
            if !seen[k]
                wall[k] = vbc; seen[k] = true
            elseif wall[k] != vbc
                constant = false
            end
        else
            constant = false
        end
    end
    return constant ? (wall, RealVector[]) : (wall, edge)
end

function staging(ctx::DeviceContext, name::Symbol, ::Type{T}, n::Int) where T
    v = get!(() -> Vector{T}(undef, n), ctx.fields, name)::Vector{T}
    length(v) == n || resize!(v, n)
    return v
end

function find_pressure_b200!(solver::PressureSolver, dt::Float64, niter::Int64 = 10;
                             boundary_velocity::Function = LagrangianVoronoi.zero_vbc,
                             rtol::Float64 = 1e-6, atol::Float64 = 1e-6, itmax::Int = 1000, # pressure.jl:219
                             krylov::Int32 = LV_SOLVER_CG)
    grid = solver.grid
    ctx = context(grid)
    n = length(grid.polygons)
    # a boundary_velocity closure is evaluated on p.edges: in the pipelined mode the edge view must have arrived first
    # (with the default zero closure the solve starts while the mesh is still on its way)
    boundary_velocity === LagrangianVoronoi.zero_vbc || sync_mesh!(grid)
    mass = staging(ctx, :mass, Float64, n); rho = staging(ctx, :rho, Float64, n); c2 = staging(ctx, :c2, Float64, n)
    P = staging(ctx, :P, Float64, n); v = staging(ctx, :v, RealVector, n)
    foreach(pin!, (mass, rho, c2, P, v))
    Threads.@threads for i in 1:n
        @inbounds begin
            p = grid.polygons[i]
            mass[i] = p.mass; rho[i] = p.rho; c2[i] = p.c2; P[i] = p.P; v[i] = p.v
        end
    end
    vbc, vbc_edge = wall_velocities(grid, boundary_velocity)
    iters = Vector{Int32}(undef, niter)
    relres = Vector{Float64}(undef, niter)
    check(ctx, ccall((:lv_find_pressure, LIB), Int32,
                     (Ptr{Cvoid}, Float64, Int32, Float64, Float64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                      Ptr{Float64}, Ptr{RealVector}, Ptr{RealVector}, Ptr{RealVector}, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}),
                     ctx.handle, dt, niter, rtol, atol, itmax, krylov, mass, rho, c2, P, v, vbc,
                     isempty(vbc_edge) ? C_NULL : pointer(vbc_edge), length(vbc_edge), P, iters,
                     solver.verbose ? pointer(relres) : C_NULL))
    Threads.@threads for i in 1:n
        @inbounds grid.polygons[i].P = P[i]      # pressure.jl:221-223
    end
    solver.verbose && @info "find_pressure!" iterations = iters relres
    return
end

# ---- mul!(y, A, x)  src/pressure.jl:119-130 (kept for API parity; uses the last assembled operator)
function mul_b200!(y::ThreadedVec{Float64}, A::PressureOperator, x::ThreadedVec{Float64}, grid::VoronoiGrid)
    ctx = context(grid)
    check(ctx, ccall((:lv_pressure_matvec, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.handle, x.val, y.val))
    return y
end

# ---- switch the package over -------------------------------------------------------------------
"""
    enable!(; device = 0)

Replace the threaded-CPU bodies of `remesh!` and `find_pressure!` by the B200 path.  Everything
that calls them (`move!`, `relaxation_step!`, every `populate_*!`, the examples) picks the new
bodies up through ordinary dispatch; signatures are unchanged.
"""
function enable!(; device::Integer = 0)
    DEVICE[] = Int32(device)
    ENABLED[] = true
    @eval LagrangianVoronoi begin
        remesh!(grid::VoronoiGrid) = $(remesh_b200!)(grid)
        find_pressure!(solver::PressureSolver, dt::Float64, niter::Int64 = 10; boundary_velocity::Function = zero_vbc) =
            $(find_pressure_b200!)(solver, dt, niter; boundary_velocity = boundary_velocity)
    end
    return
end

end # module
