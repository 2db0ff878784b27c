(cord, θ, phi, derivative, integral, u, p) -> begin
    begin
        (θ1, θ2, θ3, θ4) = (θ.depvar.u, θ.depvar.v, θ.depvar.w, θ.depvar.p)
        (phi1, phi2, phi3, phi4) = (phi[1], phi[2], phi[3], phi[4])
        let (x, y, z) = (cord[[1], :], cord[[2], :], fill(1.0, size(cord[[1], :])))
            begin
                cord1 = vcat(x, y, z)
            end
            u(cord1, θ1, phi1) .- 1.0
        end
    end
end
