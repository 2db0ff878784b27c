(cord, θ, phi, derivative, integral, u, p) -> begin
    begin
        let (t, x) = (cord[[1], :], cord[[2], :])
            begin
                cord1 = vcat(t, x)
            end
            u(cord1, θ, phi) .- (*).(-1, sin.((*).(π, x)))
        end
    end
end
