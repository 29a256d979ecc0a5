# The reference's own GPU testset (Interpolations.jl test/gpu_support.jl:4-91,115-126) with `B200Array` in place of
# `JLArray`: device results must EQUAL the CPU results (`==`, not `≈`): the kernels follow the reference operation by
# operation.  The Lanczos / Gridded testsets of that file are outside the hot path (SURVEY §8) and are left out.
using Test, Interpolations, Adapt
using B200Interpolations
const jl = b200

@testset "1d GPU Interpolation" begin
    A_x = 1.0:2.0:40.0
    A = [log(x) for x in A_x]
    itp = interpolate(A, BSpline(Cubic(Line(OnGrid()))))
    jlitp = jl(itp)
    idx = 2.0:0.17:19.0
    jlidx = jl(collect(idx))
    @test itp.(idx) == collect(jlitp.(idx)) == collect(jlitp.(jlidx))
    @test Interpolations.gradient.(Ref(itp), idx) ==
        collect(Interpolations.gradient.(Ref(jlitp), idx)) ==
        collect(Interpolations.gradient.(Ref(jlitp), jlidx))

    sitp = scale(itp, A_x)
    jlsitp = jl(sitp)
    idx = 1.0:0.4:39.0
    jlidx = jl(collect(idx))
    @test sitp.(idx) == collect(jlsitp.(idx)) == collect(jlsitp.(jlidx))
    @test Interpolations.gradient.(Ref(sitp), idx) ==
        collect(Interpolations.gradient.(Ref(jlsitp), idx)) ==
        collect(Interpolations.gradient.(Ref(jlsitp), jlidx))

    esitp = extrapolate(sitp, Flat())
    jlesitp = jl(esitp)
    idx = -1.0:0.84:41.0
    jlidx = jl(collect(idx))
    @test esitp.(idx) == collect(jlesitp.(idx)) == collect(jlesitp.(jlidx))
    @test Interpolations.gradient.(Ref(esitp), idx) ==
        collect(Interpolations.gradient.(Ref(jlesitp), idx)) ==
        collect(Interpolations.gradient.(Ref(jlesitp), jlidx))

    esitp = extrapolate(sitp, 0.0)
    jlesitp = jl(esitp)
    idx = -1.0:0.84:41.0
    jlidx = jl(collect(idx))
    @test esitp.(idx) == collect(jlesitp.(idx)) == collect(jlesitp.(jlidx))
    @test Interpolations.gradient.(Ref(esitp), idx) ==
        collect(Interpolations.gradient.(Ref(jlesitp), idx)) ==
        collect(Interpolations.gradient.(Ref(jlesitp), jlidx))
end

@testset "2d GPU Interpolation" begin
    A_x = 1.0:2.0:40.0
    A = [log(x + y) for x in A_x, y in 1.0:2.0:40.0]
    itp = interpolate(A, (BSpline(Cubic(Line(OnGrid()))), BSpline(Linear())))
    jlitp = jl(itp)
    idx = 2.0:0.17:19.0
    jlidx = jl(collect(idx))
    @test itp.(idx, idx') == collect(jlitp.(idx, idx')) == collect(jlitp.(jlidx, jlidx'))
    @test Interpolations.gradient.(Ref(itp), idx, idx') ==
        collect(Interpolations.gradient.(Ref(jlitp), idx, idx')) ==
        collect(Interpolations.gradient.(Ref(jlitp), jlidx, jlidx'))
    @test Interpolations.hessian.(Ref(itp), idx, idx') ==
        collect(Interpolations.hessian.(Ref(jlitp), idx, idx')) ==
        collect(Interpolations.hessian.(Ref(jlitp), jlidx, jlidx'))

    sitp = scale(itp, A_x, A_x)
    jlsitp = jl(sitp)
    idx = 1.0:0.4:39.0
    jlidx = jl(collect(idx))
    @test sitp.(idx, idx') == collect(jlsitp.(idx, idx')) == collect(jlsitp.(jlidx, jlidx'))
    @test Interpolations.gradient.(Ref(sitp), idx, idx') ==
        collect(Interpolations.gradient.(Ref(jlsitp), idx, idx')) ==
        collect(Interpolations.gradient.(Ref(jlsitp), jlidx, jlidx'))
    @test Interpolations.hessian.(Ref(sitp), idx, idx') ==
        collect(Interpolations.hessian.(Ref(jlsitp), idx, idx')) ==
        collect(Interpolations.hessian.(Ref(jlsitp), jlidx, jlidx'))

    esitp = extrapolate(sitp, Flat())
    jlesitp = jl(esitp)
    idx = -1.0:0.84:41.0
    jlidx = jl(collect(idx))
    @test esitp.(idx, idx') == collect(jlesitp.(idx, idx')) == collect(jlesitp.(jlidx, jlidx'))
    @test Interpolations.gradient.(Ref(esitp), idx, idx') ==
        collect(Interpolations.gradient.(Ref(jlesitp), idx, idx')) ==
        collect(Interpolations.gradient.(Ref(jlesitp), jlidx, jlidx'))

    esitp = extrapolate(sitp, 0.0)
    jlesitp = jl(esitp)
    idx = -1.0:0.84:41.0
    jlidx = jl(collect(idx))
    @test esitp.(idx, idx') == collect(jlesitp.(idx, idx')) == collect(jlesitp.(jlidx, jlidx'))
    @test Interpolations.gradient.(Ref(esitp), idx, idx') ==
        collect(Interpolations.gradient.(Ref(jlesitp), idx, idx')) ==
        collect(Interpolations.gradient.(Ref(jlesitp), jlidx, jlidx'))
end

@testset "prefilter on the device" begin
    # interpolate(::B200Array, it): copy_with_padding + A_ldiv_B_md! run on the device with the factors the reference's
    # own prefiltering_system computed on the host; coefficients agree with the CPU to the parity tolerance
    # (exactly, after B200Interpolations' strict mode: ccall :b200itp_prefilter_set_strict)
    for T in (Float32, Float64), it in (BSpline(Cubic(Periodic(OnGrid()))), BSpline(Cubic(Flat(OnCell()))),
                                        BSpline(Quadratic(Reflect(OnCell()))), BSpline(Cubic(Line(OnGrid()))))
        A = rand(T, 40, 36, 20)
        cpu = interpolate(A, it)
        gpu = interpolate(b200(A), it)
        tol = T === Float32 ? 1f-5 : 1e-12
        c, g = parent(cpu.coefs), Array(B200Interpolations.devarray(gpu.coefs))
        @test size(c) == size(g)
        @test maximum(abs.(c .- g)) <= tol * maximum(abs.(c))
        xs = T.(range(2, 19, length=1000))
        @test maximum(abs.(cpu.(xs, xs, xs) .- collect(gpu.(xs, xs, xs)))) <= 10tol * maximum(abs.(c))
    end
end

@testset "grid evaluation and bounds errors" begin
    A = rand(Float32, 30, 20)
    itp = interpolate(A, BSpline(Cubic(Line(OnGrid()))))
    jlitp = jl(itp)
    x, y = 1f0:0.25f0:30f0, 1f0:0.5f0:20f0
    @test itp(x, y) == collect(jlitp(collect(x), collect(y)))
    @test_throws BoundsError jlitp.([0.5f0, 2f0], [1f0, 1f0])
end

@testset "eltype after adaption" begin
    A_x = 1.0:2.0:40.0
    A = [log(x) for x in A_x]
    itp = interpolate(A, BSpline(Cubic(Line(OnGrid()))))
    @test eltype(adapt(Array{Float32}, itp)) === Float32
    @test eltype(adapt(Array{Float32}, scale(itp, A_x))) === Float32
    @test eltype(adapt(Array{Float32}, extrapolate(scale(itp, A_x), Flat()))) === Float32
    @test eltype(adapt(Array{Float32}, extrapolate(scale(itp, A_x), 0.0))) === Float32
    @test eltype(jl(itp)) === Float64
end
