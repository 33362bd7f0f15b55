"""Size-independent properties of the device update at the BASELINE shapes (no oracle involved): bit-reproducibility,
linearity of the actor gradient in the expert weight, learning-rate / Polyak edge cases, BC as the eps = 1 slice."""
import numpy as np
import pytest
import torch

from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import SHAPES, fill_synthetic

pytestmark = pytest.mark.gpu


def _pop(shape="ant", n=6, gemm_mode=L.GEMM_TCGEN05_BF16X3, seed=3, **hy):
    S, A, B = SHAPES[shape]
    pop = Population(PopulationSpec(n_agents=n, S=S, A=A, B=B, E=20, num_models=2, replay_capacity=4000, gemm_mode=gemm_mode))
    fill_synthetic(pop, seed=seed, **hy)
    return pop


def _state(pop):
    torch.cuda.synchronize()
    return {k: pop.t[k].clone() for k in ("actor", "actor_m", "actor_v", "q", "q_m", "q_v", "qt", "alpha", "adam_t")}


@pytest.mark.parametrize("shape", ["ant", "humanoid"])
def test_update_is_bit_reproducible(shape):
    """Fixed-order reductions, no float atomics: the same seed gives the same bits, step after step (graph replays
    included)."""
    outs = []
    for rep in range(2):
        pop = _pop(shape, n=3)
        pop.update(3, 0, True, seed=7)
        outs.append((_state(pop), pop.losses.clone()))
        pop.close()
    for k in outs[0][0]:
        assert torch.equal(outs[0][0][k], outs[1][0][k]), k
    assert torch.equal(outs[0][1], outs[1][1])


def test_actor_gradient_is_linear_in_the_expert_weight():
    """p_loss = (1 - eps) L_pi + eps MSE (SAC_expert.py:336): g(eps) = (1 - eps) g(0) + eps g(1), checked on the device
    gradients alone at the full Ant size."""
    g = {}
    for eps in (0.0, 1.0, 0.3):
        pop = _pop("ant", n=4, eps=eps)
        pop.update_phase(0, 0)          # identical critic phase (same seed-free injected state: device RNG not used)
        pop.update_phase(1, 0)
        pop.update_phase(2, 0)
        torch.cuda.synchronize()
        g[eps] = pop.debug("g_actor").clone().view(4, -1)[:, :pop.L.na].double()
        pop.close()
    mix = 0.7 * g[0.0] + 0.3 * g[1.0]
    err = (g[0.3] - mix).norm(dim=1) / mix.norm(dim=1)
    assert float(err.max()) < 2e-5, err


def test_zero_learning_rates_freeze_parameters_but_not_the_optimizer_slots():
    pop = _pop("ant", n=3, lr_q=0.0, lr_pi=0.0, lr_alpha=0.0, tau=0.0)
    before = _state(pop)
    pop.update(2, 0, True, seed=1)
    after = _state(pop)
    for k in ("actor", "q", "qt"):
        assert torch.equal(before[k], after[k]), k
    assert torch.all(after["alpha"] == torch.clamp(before["alpha"], min=1e-5))      # only the clamp acts (SAC_expert.py:348)
    assert not torch.equal(before["q_m"], after["q_m"]) and not torch.equal(before["actor_v"], after["actor_v"])
    assert torch.all(after["adam_t"] == before["adam_t"] + 2)
    pop.close()


def test_polyak_tau_one_copies_the_live_critics_bit_exactly():
    pop = _pop("ant", n=3, tau=1.0)
    pop.update(1, 0, True, seed=2)
    torch.cuda.synchronize()
    nc = pop.L.nc
    assert torch.equal(pop.t["qt"][..., :nc], pop.t["q"][..., :nc])      # target*(1-1) + live*1 in fp32 products
    pop.close()


def test_target_update_interval_gates_polyak():
    S, A, B = SHAPES["hopper"]
    pop = Population(PopulationSpec(n_agents=2, S=S, A=A, B=B, E=20, num_models=2, replay_capacity=2000, target_update_int=3,
                                    gemm_mode=L.GEMM_TCGEN05_BF16X3))
    fill_synthetic(pop, seed=5)
    qt0 = pop.t["qt"].clone()
    pop.update(1, 1, True, seed=3)          # num_timesteps = 1: 1 % 3 != 0 -> targets untouched (SAC_expert.py:475)
    torch.cuda.synchronize()
    assert torch.equal(pop.t["qt"], qt0)
    pop.update(2, 2, True, seed=3)          # steps at num_timesteps 2 and 3: the second one updates the targets
    torch.cuda.synchronize()
    assert not torch.equal(pop.t["qt"], qt0)
    pop.close()


def test_bc_equals_the_unit_expert_weight_slice_of_the_actor_phase():
    """BC._update_actor == the SAC-EO actor phase with eps = 1 (BC.py:309-363 vs SAC_expert.py:299-338)."""
    a = _pop("ant", n=3, eps=1.0)
    a.update_phase(0, 0); a.update_phase(1, 0); a.update_phase(2, 0)
    b = _pop("ant", n=3, eps=0.123)
    b.bc_update(1, use_device_rng=False)
    torch.cuda.synchronize()
    ga = a.debug("g_actor").view(3, -1)[:, :a.L.na].double()
    gb = b.debug("g_actor").view(3, -1)[:, :b.L.na].double()
    # the critic phase of `a` moved nothing the expert term depends on, so the two gradients agree to rounding
    assert float(((ga - gb).norm(dim=1) / gb.norm(dim=1)).max()) < 1e-6
    a.close(); b.close()
