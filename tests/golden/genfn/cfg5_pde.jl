(cord, θ, phi, derivative, integral, u, p) -> begin
    begin
        (θ1,) = (θ.depvar.u,)
        (phi1,) = (phi[1],)
        (a,) = (θ.p[1:1],)
        let (t, x, y, kappa) = (cord[[1], :], cord[[2], :], cord[[3], :], cord[[4], :])
            begin
                cord1 = vcat(t, x, y, kappa)
            end
            derivative(phi1, u, cord1, [[6.0554544523933395e-6, 0.0, 0.0, 0.0]], 1, θ1) .- (*).((*).(a, kappa), (+).(derivative(phi1, u, cord1, [[0.0, 0.0001220703125, 0.0, 0.0], [0.0, 0.0001220703125, 0.0, 0.0]], 2, θ1), derivative(phi1, u, cord1, [[0.0, 0.0, 0.0001220703125, 0.0], [0.0, 0.0, 0.0001220703125, 0.0]], 2, θ1)))
        end
    end
end
