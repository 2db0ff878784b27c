(cord, θ, phi, derivative, integral, u, p) -> begin
    begin
        (θ1, θ2, θ3, θ4) = (θ.depvar.u, θ.depvar.v, θ.depvar.w, θ.depvar.p)
        (phi1, phi2, phi3, phi4) = (phi[1], phi[2], phi[3], phi[4])
        let (x, y, z) = (cord[[1], :], cord[[2], :], cord[[3], :])
            begin
                cord1 = vcat(x, y, z)
                cord2 = vcat(x, y, z)
                cord3 = vcat(x, y, z)
            end
            (+).((+).(derivative(phi1, u, cord1, [[6.0554544523933395e-6, 0.0, 0.0]], 1, θ1), derivative(phi2, u, cord2, [[0.0, 6.0554544523933395e-6, 0.0]], 1, θ2)), derivative(phi3, u, cord3, [[0.0, 0.0, 6.0554544523933395e-6]], 1, θ3)) .- 0
        end
    end
end
