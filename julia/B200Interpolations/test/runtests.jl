# Acceptance tests of the Julia glue.  The cases are the ones the reference exercises for array-type-generic evaluation
# (Interpolations.jl test/gpu_support.jl: 1-D and 2-D B-splines, bare / scaled / Flat-extrapolated / filled, values,
# gradients and hessians, host and device coordinate arguments, eltype after adaption), written table-driven here: for
# every (interpolant, coordinates) pair the device result must EQUAL the CPU result (`==`, not `≈`) — the kernels follow
# the reference operation by operation.  Lanczos / Gridded interpolants are outside the hot path (SURVEY §8).
using Test, Interpolations, Adapt
using B200Interpolations
const G = Interpolations.gradient
const H = Interpolations.hessian

"CPU broadcast == device broadcast with host coordinates == device broadcast with device coordinates."
function same_everywhere(f, itp, coords...)
    ditp = b200(itp)
    dcoords = map(c -> c isa AbstractRange ? b200(collect(c)) : b200(collect(c)), coords)
    cpu = f === nothing ? itp.(coords...) : f.(Ref(itp), coords...)
    on_host_args = f === nothing ? ditp.(coords...) : f.(Ref(ditp), coords...)
    on_dev_args = f === nothing ? ditp.(dcoords...) : f.(Ref(ditp), dcoords...)
    return cpu == collect(on_host_args) == collect(on_dev_args)
end

knots = 1.0:2.0:40.0
samples1 = log.(knots)
samples2 = [log(x + y) for x in knots, y in knots]

@testset "B-spline broadcast on the B200: $(name)" for (name, base, inner, scaled_pts, wide) in (
        ("1-D cubic", interpolate(samples1, BSpline(Cubic(Line(OnGrid())))), 2.0:0.17:19.0, 1.0:0.4:39.0, -1.0:0.84:41.0),
        ("2-D cubic x linear", interpolate(samples2, (BSpline(Cubic(Line(OnGrid()))), BSpline(Linear()))), 2.0:0.17:19.0,
         1.0:0.4:39.0, -1.0:0.84:41.0))
    nd = ndims(base)
    args(r) = nd == 1 ? (r,) : (r, r')              # `itp.(idx, idx')` for the 2-D object
    sitp = nd == 1 ? scale(base, knots) : scale(base, knots, knots)
    for (obj, pts) in ((base, inner), (sitp, scaled_pts), (extrapolate(sitp, Flat()), wide), (extrapolate(sitp, 0.0), wide))
        @test same_everywhere(nothing, obj, args(pts)...)
        @test same_everywhere(G, obj, args(pts)...)
    end
    if nd == 2                                       # hessians: bare and scaled (extrapolated hessians are undefined)
        @test same_everywhere(H, base, args(inner)...)
        @test same_everywhere(H, sitp, args(scaled_pts)...)
    end
end

@testset "prefilter on the device" begin
    # interpolate(::B200Array, it): copy_with_padding + A_ldiv_B_md! run on the device with the factors the reference's
    # own prefiltering_system computed on the host; coefficients agree with the CPU to the parity tolerance, and
    # exactly in the library's strict mode (the reference's operation order, no FMA)
    strict(on) = ccall((:b200itp_prefilter_set_strict, B200Interpolations.libb200itp), Cint, (Cint,), on)
    for T in (Float32, Float64), it in (BSpline(Cubic(Periodic(OnGrid()))), BSpline(Cubic(Flat(OnCell()))),
                                        BSpline(Quadratic(Reflect(OnCell()))), BSpline(Cubic(Line(OnGrid()))))
        A = rand(T, 40, 36, 20)
        cpu = interpolate(A, it)
        gpu = interpolate(b200(A), it)
        tol = T === Float32 ? 1f-5 : 1e-12
        c, g = parent(cpu.coefs), Array(B200Interpolations.devarray(gpu.coefs))
        @test size(c) == size(g)
        @test maximum(abs.(c .- g)) <= tol * maximum(abs.(c))
        prev = strict(1)
        try
            @test parent(cpu.coefs) == Array(B200Interpolations.devarray(interpolate(b200(A), it).coefs))
        finally
            strict(prev)
        end
        xs = T.(range(2, 19, length=1000))
        @test maximum(abs.(cpu.(xs, xs, xs) .- collect(gpu.(xs, xs, xs)))) <= 10tol * maximum(abs.(c))
    end
end

@testset "grid evaluation and bounds errors" begin
    A = rand(Float32, 30, 20)
    itp = interpolate(A, BSpline(Cubic(Line(OnGrid()))))
    ditp = b200(itp)
    x, y = 1f0:0.25f0:30f0, 1f0:0.5f0:20f0
    @test itp(x, y) == collect(ditp(collect(x), collect(y)))
    @test_throws BoundsError ditp.([0.5f0, 2f0], [1f0, 1f0])
end

@testset "element types after adaption" begin
    itp = interpolate(samples1, BSpline(Cubic(Line(OnGrid()))))
    for obj in (itp, scale(itp, knots), extrapolate(scale(itp, knots), Flat()), extrapolate(scale(itp, knots), 0.0))
        @test eltype(adapt(Array{Float32}, obj)) === Float32
        @test eltype(b200(obj)) === Float64
    end
end

@testset "several devices: replicate, sharded prefilter, sharded evaluation" begin
    # needs at least two B200s; one process drives them through the library's NCCL entry points
    ndev = parse(Int, get(ENV, "B200ITP_TEST_DEVICES", "1"))
    if ndev >= 2
        comm = Comm(0:ndev-1)
        A = rand(Float32, 64, 48, 40)
        it = BSpline(Cubic(Periodic(OnGrid())))
        cpu = interpolate(A, it)
        one = interpolate(b200(A), it)
        xs = ntuple(d -> Float32.(1 .+ (size(A, d) - 1) .* rand(100_003)), 3)
        ref = collect(one.(xs...))
        @test evaluate_sharded(replicate(comm, one), xs...) == ref          # broadcast replicas
        sharded = interpolate_sharded(comm, A, it)                           # slab-sharded prefilter
        @test all(s -> Array(B200Interpolations.devarray(s.coefs)) == Array(B200Interpolations.devarray(one.coefs)), sharded)
        @test evaluate_sharded(sharded, xs...) == ref
        @test maximum(abs.(ref .- cpu.(xs...))) <= 1f-4 * maximum(abs.(A))
        @test_throws ArgumentError interpolate_sharded(comm, A, BSpline(Cubic(Flat(OnCell()))))
        close(comm)
    end
end
