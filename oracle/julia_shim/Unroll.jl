# Shim for the reference's private `Unroll` package (vm.jl:4, Common.jl:2).  TEST / BASELINE INFRASTRUCTURE.
#
# The reference uses `@unroll` on whole-slice statements, which is not the registered Unroll.jl's API:
#     @unroll wt[2:nx+1, 2:ny+1] = wn[2:nx+1, 2:ny+1] + Δt * r[2:nx+1, 2:ny+1]      (vm.jl:28)
#     @unroll data[1:nx, 1:ny] = complex(f[1:nx, 1:ny], 0.)                         (Common.jl:115)
#     @unroll rms += r[1:nx+1, 1:ny+1]^2                                            (Common.jl:235)
# i.e. "evaluate this slice statement element by element, in place".  The shim maps an assignment to a fused
# in-place broadcast over views and a `+=` to a sum of the broadcast right-hand side.
#
# NOT EXECUTED IN THE BUILD IMAGE (no julia there); bench.py only uses it when `julia` is on PATH and
# CFD_JULIA_REFERENCE points at a checkout of t-bltg/CFD_Julia.
module Unroll
export @unroll

macro unroll(ex)
  if ex isa Expr && ex.head == :(+=)
    lhs, rhs = ex.args
    return esc(:($lhs += sum(@views @. $rhs)))
  elseif ex isa Expr && ex.head == :(=)
    return esc(:(@views @. $ex))
  end
  return esc(ex)
end

end
