(cord, θ, phi, derivative, integral, u, p) -> begin
    begin
        (θ1, θ2, θ3, θ4) = (θ.depvar.u, θ.depvar.v, θ.depvar.w, θ.depvar.p)
        (phi1, phi2, phi3, phi4) = (phi[1], phi[2], phi[3], phi[4])
        let (x, y, z) = (cord[[1], :], cord[[2], :], cord[[3], :])
            begin
                cord1 = vcat(x, y, z)
                cord2 = vcat(x, y, z)
                cord3 = vcat(x, y, z)
                cord4 = vcat(x, y, z)
            end
            (+).((+).((+).((+).((*).(u(cord1, θ1, phi1), derivative(phi1, u, cord1, [[6.0554544523933395e-6, 0.0, 0.0]], 1, θ1)), (*).(u(cord2, θ2, phi2), derivative(phi1, u, cord1, [[0.0, 6.0554544523933395e-6, 0.0]], 1, θ1))), (*).(u(cord3, θ3, phi3), derivative(phi1, u, cord1, [[0.0, 0.0, 6.0554544523933395e-6]], 1, θ1))), derivative(phi4, u, cord4, [[6.0554544523933395e-6, 0.0, 0.0]], 1, θ4)), (*).(-0.01, (+).((+).(derivative(phi1, u, cord1, [[0.0001220703125, 0.0, 0.0], [0.0001220703125, 0.0, 0.0]], 2, θ1), derivative(phi1, u, cord1, [[0.0, 0.0001220703125, 0.0], [0.0, 0.0001220703125, 0.0]], 2, θ1)), derivative(phi1, u, cord1, [[0.0, 0.0, 0.0001220703125], [0.0, 0.0, 0.0001220703125]], 2, θ1)))) .- 0
        end
    end
end
