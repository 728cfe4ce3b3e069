# parity_example.jl - parity of the GPU path against the UNMODIFIED reference, to be run by a maintainer who has
# Julia >= 1.0 with the reference's packages (README.md:5-9) and a B200 with libls_cuda.so:
#
#   LS_CUDA_LIB=/path/to/libls_cuda.so LS_REFERENCE=/path/to/Fast_solver_Lippmann_Schwinger julia parity_example.jl
#
# It repeats examples/example.jl:30-91 of the reference (n = 201, Greengard_Vico operator, Duan-Rokhlin sparsifier),
# then swaps the operator and the preconditioner for their GPU twins and compares, with the tolerances of this
# repository's tests: one apply to 1e-12 (relative L2), both GMRES residual histories to 1e-8.
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia (DESIGN.md section 2: "parity unpinned").  Running it is
# what pins the oracle-based parity claims to the reference itself.
using IterativeSolvers, SpecialFunctions, SparseArrays, Distributed, SharedArrays, LinearAlgebra, FFTW, Random

const REF = get(ENV, "LS_REFERENCE", joinpath(@__DIR__, "..", "..", "Fast_solver_Lippmann_Schwinger"))
include(joinpath(REF, "src", "SparsifyingMatrix2D.jl"))      # pulls in FastConvolution.jl and Functions.jl
include(joinpath(REF, "src", "preconditioner.jl"))
include(joinpath(@__DIR__, "LSCuda.jl"))
using .LSCuda

FFTW.set_num_threads(Sys.CPU_THREADS)

# examples/example.jl:30-54
h = 0.005; k = 1 / h; a = 1
x = collect(-a/2:h:a/2); y = collect(-a/2:h:a/2)
(n, m) = length(x), length(y); N = n * m
X = repeat(x, 1, m)[:]; Y = repeat(y', n, 1)[:]
(ppw, D) = referenceValsTrapRule(); D0 = D[1]      # Duan-Rokhlin diagonal weight, used by the sparsifier (example.jl:44-45)
nu(x, y) = 0.3 * exp.(-40 * (x .^ 2 + y .^ 2)) .* (abs.(x) .< 0.48) .* (abs.(y) .< 0.48)
fastconv = buildFastConvolution(x, y, h, k, nu, quadRule = "Greengard_Vico")
As = buildSparseA(k, X, Y, D0, n, m)
Mapproxsp = As + k^2 * (buildSparseAG(k, X, Y, D0, n, m) * spdiagm(0 => nu(X, Y)))
precond = SparsifyingPreconditioner(Mapproxsp, As)

gpu_conv = GPUFastM(fastconv)
gpu_precond = GPUSparsifyingPreconditioner(Mapproxsp, As)

relerr(a, b) = norm(a - b) / norm(b)

# one apply, both modes
Random.seed!(1234)
b = randn(ComplexF64, N)
e_apply = relerr(gpu_conv * b, fastconv * b)
e_conv = relerr(FFTconvolution(gpu_conv, b), FFTconvolution(fastconv, b))
e_spmv = relerr(gpu_precond.As * b, As * b)
println("apply rel. L2 error           : ", e_apply)
println("FFTconvolution rel. L2 error  : ", e_conv)
println("As*b rel. L2 error            : ", e_spmv)

# GMRES histories, preconditioned and not (example.jl:77-93)
u_inc = exp.(k * im * X)
rhs = -k^2 * FFTconvolution(fastconv, nu(X, Y) .* u_inc)
function history(A, P)
    u = zeros(ComplexF64, N)
    info = P === nothing ? gmres!(u, A, rhs, log = true) : gmres!(u, A, rhs, Pl = P, log = true)
    return u, info[2].data[:resnorm]
end
(u_ref, h_ref) = history(fastconv, precond)
(u_gpu, h_gpu) = history(gpu_conv, gpu_precond)
(u_dev, info_dev) = gmres_gpu!(zeros(ComplexF64, N), gpu_conv, rhs, Pl = gpu_precond, log = true)
(_, h_ref0) = history(fastconv, nothing)
(_, h_gpu0) = history(gpu_conv, nothing)
hist_err(a, b) = length(a) == length(b) ? maximum(abs.(a - b) ./ b) : Inf
println("preconditioned history (IterativeSolvers on GPU operator) : ", hist_err(h_gpu, h_ref), "  (", length(h_ref), " iterations)")
println("preconditioned history (device-resident ls_gmres)          : ", hist_err(info_dev.resnorm, h_ref))
println("unpreconditioned history                                   : ", hist_err(h_gpu0, h_ref0), "  (", length(h_ref0), " iterations)")
println("solution rel. L2 difference                                : ", relerr(u_gpu, u_ref))

ok = e_apply <= 1e-12 && e_conv <= 1e-12 && e_spmv <= 1e-13 && hist_err(h_gpu, h_ref) <= 1e-8 &&
     hist_err(info_dev.resnorm, h_ref) <= 1e-8 && hist_err(h_gpu0, h_ref0) <= 1e-8
println(ok ? "PARITY OK" : "PARITY FAILED")
exit(ok ? 0 : 1)
