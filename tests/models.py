"""Model sources shared by the tests: the reference's examples / benchmark models verbatim
(examples/1D_ssm.jl:7-16, 2D_ssm.jl:7-17, linear_regression.jl:17-27, eight_schools.jl:7-17,
benchmarks/ssm/WeightedSampling/lgssm1d.jl:18-24)."""

SSM1D = '''
@model function ssm(obs)
    x{1} .= 0.0
    v .= 0.0
    for (t, o) in enumerate(obs)
        x{t + 1} .= x{t} + v
        dv ~ Normal(0.0, 0.1)
        v .= v + dv
        o => Normal(x{t + 1}, 1.0)
    end
end
'''

SSM2D = '''
@model function ssm(obs)
    I2 = [1.0 0.0; 0.0 1.0]
    x{1} .= [0.0, 0.0]
    v .= [1.0, 0.0]
    for (t, o) in enumerate(obs)
        x{t + 1} .= x{t} + v
        dv ~ MvNormal([0.0, 0.0], 0.1 * I2)
        v .= v + dv
        o => MvNormal(x{t + 1}, 0.5 * I2)
    end
end
'''

SSM2D_FILTER = '''
@model function ssm2d_filter(obs)
    I2 = [1.0 0.0; 0.0 1.0]
    x .= [0.0, 0.0]
    v .= [1.0, 0.0]
    for o in obs
        x .= x + v
        dv ~ MvNormal([0.0, 0.0], 0.1 * I2)
        v .= v + dv
        o => MvNormal(x, 0.5 * I2)
    end
end
'''

LINREG = '''
@model function linear_regression(xs, ys)
    α ~ Normal(0.0, 10.0)
    β ~ Normal(0.0, 10.0)
    for (x, y) in zip(xs, ys)
        y => Normal(α + β * x, 1.0)
        if resampled
            α << autoRW()
            β << autoRW()
        end
    end
end
'''

SCHOOLS = '''
@model function eight_schools(J, y, σ)
    μ ~ Normal(0.0, 5.0)
    τ ~ Exponential(5.0)
    θ .= zeros(J)
    for j in 1:J
        θ[j] ~ Normal(μ, τ)
        y[j] => Normal(θ[j], σ[j])
        μ << autoRW(; diversity=0.9)
        τ << autoRW(1e-3, (0.0, Inf); diversity=0.9)
    end
end
'''

LGSSM1D = '''
@model function lgssm1d(data, a, q, r, x0_std)
    x ~ Normal(0.0, x0_std)
    for y in data
        x ~ Normal(a * x, q)
        y => Normal(x, r)
    end
end
'''

SCHOOLS_Y = [28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0]
SCHOOLS_SIGMA = [15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0]

# benchmarks/multilevel/WeightedSampling/model.jl:20-41, verbatim
HIER = '''
@model function hierarchical_regression(J, groups)
    mu_alpha ~ Normal(0.0, 10.0)
    tau_alpha ~ Exponential(1.0)
    beta ~ Normal(0.0, 10.0)
    sigma ~ Exponential(1.0)
    for j in 1:J
        alpha{j} ~ Normal(mu_alpha, tau_alpha)
        obs = groups[j]
        for (x, y) in obs
            y => Normal(alpha{j} + beta * x, sigma)
            if resampled
                alpha{j} << autoRW(diversity = 0.1)
            end
        end      
        if j % 10 == 0
            mu_alpha << autoRW(diversity = 0.1)
            tau_alpha << autoRW(1e-3, (0.0, Inf); diversity = 0.1)
            beta << autoRW(diversity = 0.1)
            sigma << autoRW(1e-3, (0.0, Inf); diversity = 0.1)
        end
    end
end
'''


def simulate_hier(J, n_obs, seed=42, mu_alpha=5.0, tau_alpha=2.0, beta=3.0, sigma=1.0):
    """benchmarks/multilevel/simulate.jl:21-38 (same generative process; NumPy stream instead of Julia's)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    alpha = mu_alpha + tau_alpha * rng.standard_normal(J)
    groups = []
    for j in range(J):
        xs = rng.standard_normal(n_obs)
        ys = alpha[j] + beta * xs + sigma * rng.standard_normal(n_obs)
        groups.append([(float(a), float(b)) for a, b in zip(xs, ys)])
    return groups, alpha
