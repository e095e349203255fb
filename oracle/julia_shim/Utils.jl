# Shim for the reference's private `Utils` package (vm.jl:5): only `boolenv` is used by the scripts
# (e.g. lid_driven_cavity.jl:193 `if boolenv("BENCH")`).  TEST / BASELINE INFRASTRUCTURE, not executed in the build image.
module Utils
export boolenv

boolenv(key::AbstractString) = lowercase(get(ENV, key, "")) in ("1", "true", "yes", "on")

end
