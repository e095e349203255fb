# Times the REFERENCE's own `numerical` (19_NS2D_Vortex_Merger/vm.jl:12-90, included unmodified) the way
# vm.jl:138 does (`@timed`), on the vortex-merger initial condition of vm.jl:107-128.
#
#   CFD_JULIA_REFERENCE=/path/to/CFD_Julia julia -t 1 oracle/julia_shim/time_vm.jl <n> <nsteps> <dt>
#
# Prints one JSON line: {"n":…, "steps":…, "seconds":…, "pt_steps_per_s":…, "julia_threads":…, "fftw_threads":…}.
# Needs FFTW.jl and BenchmarkTools.jl in the active Julia environment; Unroll / Utils come from this directory.
# NOT EXECUTED IN THE BUILD IMAGE (no julia there) -- see bench.py `julia_probe`.
push!(LOAD_PATH, @__DIR__)
const REF = get(ENV, "CFD_JULIA_REFERENCE", "")
isdir(REF) || error("CFD_JULIA_REFERENCE does not point at a CFD_Julia checkout")
cd(mktempdir())                                                  # vm.jl writes vm<m>.txt into the working directory
include(joinpath(REF, "19_NS2D_Vortex_Merger", "vm.jl"))         # defines numerical / main; main() does not run on include
import FFTW

function run(n::Int, nt::Int, Δt::Float64)
  Δx = 2π / n
  x = [Δx * (i - 1) for i ∈ 1:n+1]
  wn = Array{Float64}(undef, n + 2, n + 2)
  vm_ic(n, n, x, x, wn)
  wn[1, :] = wn[n+1, :]; wn[:, 1] = wn[:, n+1]; wn[n+2, :] = wn[2, :]; wn[:, n+2] = wn[:, 2]
  w1 = copy(wn)
  numerical(n, n, 1, Δx, Δx, Δt, 1000., x, x, w1, 1)              # compile + FFTW plan warm-up
  t = @elapsed numerical(n, n, nt, Δx, Δx, Δt, 1000., x, x, wn, 1)
  println("{\"n\":$n,\"steps\":$nt,\"seconds\":$t,\"pt_steps_per_s\":$(n * n * nt / t),",
          "\"julia_threads\":$(Threads.nthreads()),\"fftw_threads\":$(FFTW.get_num_threads())}")
end

run(parse(Int, ARGS[1]), parse(Int, ARGS[2]), parse(Float64, ARGS[3]))
