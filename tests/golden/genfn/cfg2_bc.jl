(cord, θ, phi, derivative, integral, u, p) -> begin
    begin
        let (x, y) = (cord[[1], :], cord[[2], :])
            begin
                cord1 = vcat(x, y)
            end
            u(cord1, θ, phi) .- 0.0
        end
    end
end
