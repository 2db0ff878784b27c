(cord, θ, phi, derivative, integral, u, p) -> begin
    begin
        let (t, x) = (cord[[1], :], cord[[2], :])
            begin
                cord1 = vcat(t, x)
            end
            (+).((+).(derivative(phi, u, cord1, [[6.0554544523933395e-6, 0.0]], 1, θ), (*).(u(cord1, θ, phi), derivative(phi, u, cord1, [[0.0, 6.0554544523933395e-6]], 1, θ))), (*).(-0.003183098861837907, derivative(phi, u, cord1, [[0.0, 0.0001220703125], [0.0, 0.0001220703125]], 2, θ))) .- 0
        end
    end
end
