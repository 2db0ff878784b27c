(cord, θ, phi, derivative, integral, u, p) -> begin
    begin
        let (x, y) = (cord[[1], :], cord[[2], :])
            begin
                cord1 = vcat(x, y)
            end
            (+).(derivative(phi, u, cord1, [[0.0001220703125, 0.0], [0.0001220703125, 0.0]], 2, θ), derivative(phi, u, cord1, [[0.0, 0.0001220703125], [0.0, 0.0001220703125]], 2, θ)) .- (*).((*).(-1, sin.((*).(π, x))), sin.((*).(π, y)))
        end
    end
end
