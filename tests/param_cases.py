"""Non-default parameter sets for the parametrised parity tests (the config scripts are parameter surfaces, not constants).
Each case mutates the abx_sim_config BOTH sides are built from: the oracle (OracleSim.from_config, seed cascade of the config scripts) and the
product (CUDA library on the GPU box, its CPU emulation in the CPU suite)."""

NS = 10 ** 9


def _zi_small_population(c):          # other agent counts per strategy group, other surplus ranges / eta
    counts, rmin, rmax, eta = [30, 25, 20, 15, 10, 6, 4], [0, 10, 0, 100, 0, 250, 50], [300, 500, 800, 1500, 2500, 400, 700], [1.0, 0.9, 0.8, 1.0, 0.7, 0.8, 1.0]
    for g in range(7):
        c.groups[g].count, c.groups[g].r_min, c.groups[g].r_max, c.groups[g].eta = counts[g], rmin[g], rmax[g], eta[g]
    c.n_agents = 1 + sum(counts)


def _zi_fast_arrivals(c):             # lambda_a x 2, fundamental kappa / volatility, observation noise, megashocks 10 x as frequent
    c.lambda_a = 2e-12
    c.kappa = 3.3e-12
    c.fund_vol = 2e-4
    c.sigma_n = 250000.0
    c.megashock_lambda_a = 2.77778e-12
    c.megashock_mean = 500.0


def _zi_small_qmax(c):                # q_max 4 (theta has 8 entries; position limits bind), order size 50, other cash, 3 noise values, 2 s delay
    c.q_max = 4
    c.order_size = 50
    c.starting_cash = 5000000
    c.n_noise = 3
    c.default_computation_delay_ns = 2 * NS
    c.latency_hi = 5000000.0


def _zi100_jitter(c):                 # cubic latency model: other jitter / clip / unit, pipeline + exchange computation delay
    c.jitter = 0.5
    c.jitter_clip = 0.1
    c.jitter_unit = 10.0
    c.exchange_pipeline_delay_ns = 40000
    c.exchange_computation_delay_ns = 1000
    c.agent_kappa = 1.0e-14


def _r3_counts(c):                    # other population mix: 30 noise, 20 value, 3 momentum agents; sizes 5..15
    c.n_noise_agents, c.n_value_agents, c.n_momentum_agents = 30, 20, 3
    c.n_agents = 1 + 30 + 20 + c.n_mm_agents + 3 + c.n_pov_exec
    c.size_lo, c.size_hi = 5, 15
    c.mom_min_size, c.mom_max_size = 2, 7


def _r3_market_maker(c):              # market maker: pov 0.2, 10 ticks, window 3, min size 7, wakes every 2 s; value agents 30 % aggressive, depth 3
    c.mm_pov = 0.2
    c.mm_num_ticks = 10
    c.mm_window_size = 3
    c.mm_min_order_size = 7
    c.mm_wake_ns = 2 * NS
    c.value_percent_aggr = 0.3
    c.value_depth_spread = 3
    c.lambda_a = 1.5e-10
    c.mom_wake_ns = 7 * NS


SPARSE_ZI_CASES = [(1000, _zi_small_population, 4242), (1000, _zi_fast_arrivals, 17), (1000, _zi_small_qmax, 99), (100, _zi100_jitter, 123456789), (100, _zi_small_qmax, 5)]
RMSC03_CASES = [(False, _r3_counts, 1001), (False, _r3_market_maker, 7), (True, _r3_market_maker, 123456789)]
