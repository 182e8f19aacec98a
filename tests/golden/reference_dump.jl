# reference_dump.jl -- run the UNMODIFIED reference (Gabisanth/MaximumAreaCoverageOptimization.jl) on the inputs of
# tests/golden/reference_inputs.txt and write tests/golden/reference_outputs.txt, so that the oracle (oracle/) and
# the CUDA path are pinned by the reference's own arithmetic instead of by a restatement of it.
#
#     julia tests/golden/reference_dump.jl /path/to/MaximumAreaCoverageOptimization.jl
#     python -m pytest tests/test_reference_outputs.py            # oracle vs the dump (CPU)
#     python -m pytest tests/test_reference_outputs.py -m gpu     # CUDA path vs the dump (B200)
#
# NOT EXECUTED in this repository's build container or on its GPU boxes: neither has Julia.  Until somebody runs it,
# parity stays "unpinned by the reference" (DESIGN.md section 2) and tests/test_reference_outputs.py skips.
#
# What is called, all of it straight from the reference's source files (nothing is redefined here):
#   AreaCoverageCalculation.createPOI / make_circles / calculateArea / rmvCoveredPOI   src/AreaCoverageCalculation.jl:11-137
#   createObjective(cells, N, r_max) -> AreaMaxObjective(x)                              src/TDM_Constraints.jl:33-51
#       (byte-identical twin of src/TDM_STATIC_opt.jl:82-100, which additionally needs DirectSearch.jl to load)
#   create_cons3(pre, FOV, d_lim), cons7, cons8, cons1/2/3_progressive                   src/TDM_Constraints.jl:54-221
# cons7 / cons8 / cons*_progressive read the Main-scope globals N, FOV, r_max, exactly as src/FullSimulation.jl sets them.
#
# src/Base_Functions.jl (included by AreaCoverageCalculation.jl) says `using LinearAlgebra, Rotations, Random,
# RobotDynamics, Plots` for its plotting / ALTRO helpers, none of which this path calls.  If one of those packages is
# not installed, an EMPTY stand-in package of that name is put on the LOAD_PATH so that the file loads unmodified.
#
# Output format (one line per item, Float64 as the 16 hex digits of its bit pattern):
#   tan_half_fov <hex>                 Julia's tan(FOV/2), FOV = 100/180*pi  (the value the comparisons must use for cons3/cons7)
#   case <name> P <entries>
#   y <obj> <area> <cons3 0|1> <cons7 0|1> <cons8 0|1> <cons1_progressive> <cons2_progressive> <cons3_progressive>   (B lines)
#   rmv <b> <remaining entries> <sum of the 1-based indices of the removed entries>       (every 50th candidate)

const REF = length(ARGS) >= 1 ? ARGS[1] : get(ENV, "COVERAGE_REFERENCE", "/root/reference")
const HERE = @__DIR__

# ---- stand-ins for optional plotting / dynamics packages that this path never calls ----
let stub_dir = mktempdir()
    for pkg in ("Rotations", "RobotDynamics", "Plots")
        if Base.find_package(pkg) === nothing
            mkpath(joinpath(stub_dir, pkg, "src"))
            write(joinpath(stub_dir, pkg, "src", pkg * ".jl"), "module $pkg\nend\n")
            @info "package $pkg is not installed: using an empty stand-in (the coverage path does not call it)"
        end
    end
    push!(LOAD_PATH, stub_dir)
end

include(joinpath(REF, "src", "TDM_Constraints.jl"))   # includes AreaCoverageCalculation.jl -> Base_Functions.jl

bits2f(s::AbstractString) = reinterpret(Float64, parse(UInt64, s; base = 16))
f2bits(v::Real) = string(reinterpret(UInt64, Float64(v)); base = 16, pad = 16)

struct DumpCells                       # what createObjective reads: cells.points_of_interest
    points_of_interest::Vector{Vector{Float64}}
end

mutable struct Case
    name::String
    N::Int
    r_max::Vector{Float64}
    pre::Vector{Float64}
    d_lim::Vector{Float64}
    points::Vector{Vector{Float64}}
    X::Vector{Vector{Float64}}
end

function read_cases(path)
    cases = Case[]
    for line in eachline(path)
        (isempty(line) || startswith(line, "#")) && continue
        tok = split(line)
        if tok[1] == "case"
            push!(cases, Case(tok[2], parse(Int, tok[4]), Float64[], Float64[], Float64[], Vector{Float64}[], Vector{Float64}[]))
        elseif tok[1] == "r_max"
            cases[end].r_max = bits2f.(tok[2:end])
        elseif tok[1] == "pre"
            cases[end].pre = bits2f.(tok[2:end])
        elseif tok[1] == "d_lim"
            cases[end].d_lim = bits2f.(tok[2:end])
        elseif tok[1] == "point"
            push!(cases[end].points, bits2f.(tok[2:end]))
        elseif tok[1] == "x"
            push!(cases[end].X, bits2f.(tok[2:end]))
        end
    end
    return cases
end

# globals the reference's constraint functions read from Main scope (src/FullSimulation.jl:733-757)
global FOV = 100 / 180 * pi
global N = 5
global r_max = Float64[]

function main()
    cases = read_cases(joinpath(HERE, "reference_inputs.txt"))
    open(joinpath(HERE, "reference_outputs.txt"), "w") do io
        println(io, "# written by tests/golden/reference_dump.jl from the reference at ", REF, " with Julia ", VERSION)
        println(io, "tan_half_fov ", f2bits(tan(FOV / 2)))
        for c in cases
            global N = c.N
            global r_max = c.r_max
            points = isempty(c.points) ? AreaCoverageCalculation.createPOI(5.0, 5.0, 100.0, 100.0) : c.points
            cells = DumpCells(points)
            objective = createObjective(cells, c.N, c.r_max)
            pre_circles = AreaCoverageCalculation.make_circles(c.pre)
            cons3 = create_cons3(pre_circles, FOV, c.d_lim)
            println(io, "case ", c.name, " P ", length(points))
            for x in c.X
                area = AreaCoverageCalculation.calculateArea(x, points)
                println(io, "y ", f2bits(objective(x)), " ", f2bits(area), " ", Int(cons3(x)), " ", Int(cons7(x)), " ",
                        Int(cons8(x)), " ", f2bits(cons1_progressive(x)), " ", f2bits(cons2_progressive(x)), " ",
                        f2bits(cons3_progressive(x)))
            end
            for b in 1:50:length(c.X)
                pts = deepcopy(points)
                keys = [(p[1], p[2]) for p in pts]
                left = AreaCoverageCalculation.rmvCoveredPOI(c.X[b], pts)
                # list order is kept, so the removed entries are found by walking both lists
                removed_sum = 0
                k = 1
                for (idx, key) in enumerate(keys)
                    if k <= length(left) && (left[k][1], left[k][2]) == key
                        k += 1
                    else
                        removed_sum += idx
                    end
                end
                println(io, "rmv ", b - 1, " ", length(left), " ", removed_sum)
            end
        end
    end
    println("wrote ", joinpath(HERE, "reference_outputs.txt"))
end

main()
