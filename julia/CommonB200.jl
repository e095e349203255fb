# CommonB200.jl -- drop-in replacements for the vortex-merger functions of CFD_Julia's Common.jl / vm.jl / tgv.jl,
# implemented by ccall into libvmk.so (hand-written sm_100a CUDA kernels behind the C ABI of include/vmk.h).
#
#   include("CommonB200.jl"); using .CommonB200
#   fps(nx, ny, Δx, Δy, u, e, data, data1, f, s, ε)          replaces Common.jl:97-125
#   vm_rhs(nx, ny, Δx, Δy, re, w, u, e, data, data1, r, s, f) replaces Common.jl:132-182
#   numerical(nx, ny, nt, Δx, Δy, Δt, re, x, y, wn, ns)       replaces 19_NS2D_Vortex_Merger/vm.jl:12-90
#   numerical(nx, ny, nt, Δx, Δy, Δt, re, wn)                 replaces 19_NS2D_Vortex_Merger/tgv.jl:13-79
#   ps_fft(nx, ny, Δx, Δy, f, ε)                              replaces 12_Poisson_Solver_FFT/fft_p.jl:8-42
#
# Same positional arguments, same in-place mutation, same return values, Array{Float64,2} throughout.  The dead
# scratch arguments u, e, data, data1 are accepted and ignored (the reference never reads them either: `e` is
# rebound at Common.jl:117 and `u` is untouched).  The library path comes from ENV["VMK_LIB"] or the repo layout.
#
# Julia is not installed in the build container, so this file is exercised only where a Julia binary exists;
# the same ABI is exercised in CI through ctypes (cfd_julia_b200/common.py), which this file mirrors line by line.
module CommonB200

export fps, vm_rhs, numerical, numerical_hybrid, numerical_ps23, numerical_ps32, numerical_ldc, write_field, read_field, ps_fft
export vmk_plan, vmk_upload, vmk_step, vmk_download, vmk_plans_multi, numerical_multi

const libvmk = get(ENV, "VMK_LIB", joinpath(@__DIR__, "..", "cfd_julia_b200", "libvmk.so"))

mutable struct Plan
  handle::Ptr{Cvoid}
end

const PLANS = Dict{Tuple{Int,Int},Plan}()

lasterror() = unsafe_string(ccall((:vmk_last_error, libvmk), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : error("vmk error $rc: $(lasterror())")

# the reference API has no plan object: cache one per grid size (device buffers, tables, stream, CUDA graph)
vmk_plan(nx::Integer, ny::Integer) = get!(PLANS, (Int(nx), Int(ny))) do
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:vmk_plan_create, libvmk), Cint, (Int64, Int64, Ptr{Ptr{Cvoid}}), nx, ny, h))
  p = Plan(h[])
  finalizer(q -> ccall((:vmk_plan_destroy, libvmk), Cint, (Ptr{Cvoid},), q.handle), p)
  p
end

ghosted(a, nx, ny, name) = size(a) == (nx + 2, ny + 2) || throw(BoundsError(a, (nx + 2, ny + 2)))

fps(nx, ny, Δx, Δy, u, e, data, data1, f::Matrix{Float64}, s::Matrix{Float64}, ε=1.e-6) = begin
  size(f) == (nx, ny) || throw(BoundsError(f, (nx, ny)))
  ghosted(s, nx, ny, "s")
  check(ccall((:vmk_fps, libvmk), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
              vmk_plan(nx, ny).handle, Δx, Δy, f, s, ε))
  return
end

ps_fft(nx, ny, Δx, Δy, f::Matrix{Float64}, ε) = begin
  size(f) == (nx + 1, ny + 1) || throw(BoundsError(f, (nx + 1, ny + 1)))
  u = Array{Float64}(undef, nx, ny)
  check(ccall((:vmk_ps_fft, libvmk), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
              vmk_plan(nx, ny).handle, Δx, Δy, f, u, ε))
  return u
end

vm_rhs(nx, ny, Δx, Δy, re, w::Matrix{Float64}, u, e, data, data1, r::Matrix{Float64}, s::Matrix{Float64},
       f::Matrix{Float64}) = begin
  ghosted(w, nx, ny, "w"); ghosted(r, nx, ny, "r"); ghosted(s, nx, ny, "s")
  size(f) == (nx, ny) || throw(BoundsError(f, (nx, ny)))
  check(ccall((:vmk_rhs, libvmk), Cint,
              (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
              vmk_plan(nx, ny).handle, Δx, Δy, re, w, r, s, f))
  return
end

# snapshot trampoline: the C side brings wn to the host and calls back with the step number
# An exception must not unwind through the C frames of the library (undefined behaviour for ccall): the first one is
# kept, later callbacks do nothing, and it is rethrown after the call has returned (same contract as common.py's _snap).
const SNAP_ERROR = Ref{Any}(nothing)
function snap_trampoline(k::Int64, wn::Ptr{Cdouble}, user::Ptr{Cvoid})::Cvoid
  SNAP_ERROR[] === nothing || return
  f = unsafe_pointer_to_objref(user)::Base.RefValue{Function}
  try
    f[](k)
  catch err
    SNAP_ERROR[] = err
  end
  return
end
function rethrow_snapshot_error()
  err = SNAP_ERROR[]
  SNAP_ERROR[] = nothing
  err === nothing || throw(err)
  return
end

# vm.jl flavour: writes "vm<m>.txt" every nt ÷ ns steps like vm.jl:78-86 (the record index is incremented here; the
# reference forgets to and overwrites vm1.txt ten times, compare hybrid.jl:81)
numerical(nx, ny, nt, Δx, Δy, Δt, re, x, y, wn::Matrix{Float64}, ns) = begin
  ghosted(wn, nx, ny, "wn")
  freq = nt ÷ ns
  m = Ref(1)
  cb = Ref{Function}(k -> begin
    @show k
    write_field("vm$(m[]).txt", x, y, wn[2:nx+2, 2:ny+2])  # vm.jl:81-85, through the library's writer
    m[] += 1
  end)
  out = Array{Float64}(undef, nx + 1, ny + 1)
  GC.@preserve cb begin
    check(ccall((:vmk_numerical, libvmk), Cint,
                (Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cvoid},
                 Ptr{Cvoid}),
                vmk_plan(nx, ny).handle, nt, Δx, Δy, Δt, re, wn, out, freq,
                @cfunction(snap_trampoline, Cvoid, (Int64, Ptr{Cdouble}, Ptr{Cvoid})), pointer_from_objref(cb)))
  end
  rethrow_snapshot_error()
  return out
end

# tgv.jl flavour: no snapshots
numerical(nx, ny, nt, Δx, Δy, Δt, re, wn::Matrix{Float64}) = begin
  ghosted(wn, nx, ny, "wn")
  out = Array{Float64}(undef, nx + 1, ny + 1)
  check(ccall((:vmk_numerical, libvmk), Cint,
              (Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cvoid},
               Ptr{Cvoid}),
              vmk_plan(nx, ny).handle, nt, Δx, Δy, Δt, re, wn, out, 0, C_NULL, C_NULL))
  return out
end

# hybrid.jl flavour (20_NS2D_Hybrid_Solver/hybrid.jl:14-90): RK3 / Crank-Nicolson in Fourier space, and the pseudo-spectral
# solvers with the 2/3 rule (22_NS2D_PseudoSpectral_23_Rule/pseudospectral_23_rule.jl:13-89) and the 3/2 rule
# (21_NS2D_PseudoSpectral_32_Rule/pseudospectral_32_rule.jl:13-89): same time loop, spectral Jacobian.  Same arguments as vm.jl's numerical; wn is only read; returns ut = real(ifft(wf)) with the periodic
# duplicates, (nx+1) x (ny+1), and writes "vm<m>.txt" every nt ÷ ns steps (hybrid.jl:71-86) through the library's writer
# (byte-identical to the scripts' `write(io, "$(x[i]) $(y[j]) $(ut[i, j])\n")` loop, minutes faster at 8192^2).
for (jlname, csym) in ((:numerical_hybrid, :vmk_hybrid_numerical), (:numerical_ps23, :vmk_ps23_numerical),
                       (:numerical_ps32, :vmk_ps32_numerical))
  @eval $jlname(nx, ny, nt, Δx, Δy, Δt, re, x, y, wn::Matrix{Float64}, ns) = begin
    ghosted(wn, nx, ny, "wn")
    freq = nt ÷ ns
    ut = Array{Float64}(undef, nx + 1, ny + 1)
    m = Ref(1)
    cb = Ref{Function}(k -> begin
      @show k
      write_field("vm$(m[]).txt", x, y, ut)
      m[] += 1
    end)
    GC.@preserve cb begin
      check(ccall(($(QuoteNode(csym)), libvmk), Cint,
                  (Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cvoid},
                   Ptr{Cvoid}),
                  vmk_plan(nx, ny).handle, nt, Δx, Δy, Δt, re, wn, ut, freq,
                  @cfunction(snap_trampoline, Cvoid, (Int64, Ptr{Cdouble}, Ptr{Cvoid})), pointer_from_objref(cb)))
    end
    rethrow_snapshot_error()
    return ut
  end
end

# the scripts' text dump (vm.jl:81-85,132-136,142-146) and plotting.jl:14-28's reader, native code in the library
write_field(path::AbstractString, x::Vector{Float64}, y::Vector{Float64}, ut::Matrix{Float64}) = begin
  nx1, ny1 = size(ut)
  (length(x) >= nx1 && length(y) >= ny1) || throw(BoundsError(x, nx1))
  check(ccall((:vmk_write_field, libvmk), Cint, (Cstring, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int64),
              path, x, y, ut, nx1, ny1))
  return
end

read_field(path::AbstractString, nx::Integer, ny::Integer) = begin
  n = (nx + 1) * (ny + 1)
  x, y, w = zeros(n), zeros(n), zeros(n)
  rows = Ref{Int64}(0)
  check(ccall((:vmk_read_field, libvmk), Cint, (Cstring, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Int64}),
              path, x, y, w, n, rows))
  rows[] == n || throw(DimensionMismatch("$path: $(rows[]) rows, expected $n"))
  return x[1:nx+1], reshape(y, (nx + 1, ny + 1))[1, :], reshape(w, (nx + 1, ny + 1))
end

# lid_driven_cavity.jl flavour (18_NS2D_Lid_Driven_Cavity/lid_driven_cavity.jl:59-117): same arguments, wn, sn
# ((nx+1) x (ny+1) node arrays) and rms are mutated in place.  The plan has size 2nx x 2ny (sine transform = periodic
# transform of the odd extension).
numerical_ldc(nx, ny, nt, Δx, Δy, Δt, re, wn::Matrix{Float64}, sn::Matrix{Float64}, rms::Vector{Float64}) = begin
  size(wn) == (nx + 1, ny + 1) || throw(BoundsError(wn, (nx + 1, ny + 1)))
  size(sn) == (nx + 1, ny + 1) || throw(BoundsError(sn, (nx + 1, ny + 1)))
  length(rms) >= nt || throw(BoundsError(rms, nt))
  check(ccall((:vmk_ldc_numerical, libvmk), Cint,
              (Ptr{Cvoid}, Int64, Int64, Int64, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
              vmk_plan(2nx, 2ny).handle, nx, ny, nt, Δx, Δy, Δt, re, wn, sn, rms))
  return
end

# ---- several GPUs driven by this one Julia process (slab decomposition along j) ---------------------------------------
# plans = vmk_plans_multi(nx, ny, ngpu); numerical_multi(plans, nt, Δx, Δy, Δt, re, wn) steps all ranks concurrently
# (vmk_step is asynchronous) and gathers every rank's rows back into wn.
vmk_plans_multi(nx::Integer, ny::Integer, ngpu::Integer) = begin
  plans = Plan[]
  for r in 0:ngpu-1
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:vmk_plan_create_on, libvmk), Cint, (Cint, Int64, Int64, Cint, Cint, Ptr{Ptr{Cvoid}}), r, nx, ny, r, ngpu, h))
    p = Plan(h[])
    finalizer(q -> ccall((:vmk_plan_destroy, libvmk), Cint, (Ptr{Cvoid},), q.handle), p)
    push!(plans, p)
  end
  handles = [p.handle for p in plans]
  for p in plans
    check(ccall((:vmk_peer_attach_local, libvmk), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), p.handle, handles))
  end
  plans
end

numerical_multi(plans::Vector{Plan}, nt, Δx, Δy, Δt, re, wn::Matrix{Float64}) = begin
  for p in plans; vmk_upload(p, wn); end                 # every rank takes its own columns of wn
  for p in plans; vmk_step(p, Δx, Δy, Δt, re, nt); end    # asynchronous: the ranks run concurrently
  for p in plans; vmk_download(p, wn); end               # every rank writes its own columns (and ghosts) back
  nx, ny = size(wn, 1) - 2, size(wn, 2) - 2
  return wn[2:nx+2, 2:ny+2]
end

# device-resident pieces, for callers that want to keep the field on the GPU between calls
vmk_upload(p::Plan, wn::Matrix{Float64}) = check(ccall((:vmk_upload, libvmk), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), p.handle, wn))
vmk_step(p::Plan, Δx, Δy, Δt, re, nsteps) =
  check(ccall((:vmk_step, libvmk), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cdouble, Int64), p.handle, Δx, Δy, Δt, re, nsteps))
vmk_download(p::Plan, wn::Matrix{Float64}, psi::Union{Matrix{Float64},Nothing}=nothing) =
  check(ccall((:vmk_download, libvmk), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), p.handle, wn,
              psi === nothing ? C_NULL : pointer(psi)))

end # module
