#!/usr/bin/env python
"""Instruction-count model of lidf_kernel's warp task queue (CPU only, planning tool).

Per (sample, angle) task the number of exact steps (stage A) and polynomial steps (stage B) is taken
from a NumPy run of the two-stage iteration with the kernel's hand-over rule; a warp then replays its
768 tasks through the two queues exactly like warp_lidf (steps per bookkeeping round, batched
hand-out) and the issued warp instructions are counted with per-phase costs read off the SASS.
Used to rank scheduling policies (round length, hand-out batch, hand-out order) before touching
the kernel; the model reproduces the measured optimum of the round length (20 steps).
usage: python tools/lidf_queue_sim.py [warps]"""
import sys

import numpy as np

TAU, SPW = 1.6e-2, 64
# warp instructions per unit (all pipes), from cuobjdump of lidf_kernel
C_A_STEP, C_B_STEP, C_ROUND, C_HAND_A, C_HAND_B = 46.0, 10.0, 14.0, 30.0, 105.0


def task_steps(a, b):
    """(nA, nB, direct) per task [n, 12]: exact steps until hand-over / convergence, polynomial steps."""
    n = a.size
    rd = np.pi / 180
    thetas = [10 * i for i in range(1, 9)] + [82, 84, 86, 88]
    nA = np.zeros((n, 12), int)
    nB = np.zeros((n, 12), int)
    direct = np.zeros((n, 12), bool)
    for i, th in enumerate(thetas):
        t2 = np.full(n, 2 * rd * th)
        x = t2.copy()
        stage = np.zeros(n, int)              # 0 running A, 1 handed over, 2 converged in A
        while (stage == 0).any():
            act = stage == 0
            s, c = np.sin(x), np.cos(x)
            dx = 0.5 * (s * (a + b * c) - x + t2)
            yp = a * c + b * (2 * c * c - 1)
            done = ~(np.abs(dx) > 1e-8)
            sw = (np.abs(dx) <= TAU * 0.5 * (1 - yp)) & ~done
            x = np.where(act, x + dx, x)
            nA[:, i] += act
            stage = np.where(act & done, 2, np.where(act & sw, 1, stage))
        direct[:, i] = stage == 2
        act = stage == 1
        while act.any():                      # exact map instead of its Taylor model: same step count
            dx = 0.5 * (a * np.sin(x) + 0.5 * b * np.sin(2 * x) - x + t2)
            x = np.where(act, x + dx, x)
            nB[:, i] += act
            act &= np.abs(dx) > 1e-8
    return nA, nB, direct


def replay(steps, per_round, batch, c_step, c_hand, order=None):
    """One queue of warp_lidf: returns issued warp instructions and the lane-step utilisation."""
    steps = np.asarray(steps)
    order = np.arange(steps.size) if order is None else order
    nxt = 32
    left = steps[order[:32]].astype(int).copy()
    cost = useful = 0.0
    total = steps.size
    while True:
        cost += per_round * c_step + C_ROUND
        useful += np.minimum(left, per_round).sum()
        left = np.maximum(left - per_round, 0)
        idle = left == 0
        if nxt < total:
            if idle.sum() >= batch:
                k = int(idle.sum())
                new = order[nxt:nxt + k]
                fill = np.zeros(k, int)
                fill[:new.size] = steps[new]
                left[idle] = fill
                nxt += k
                cost += c_hand
        elif idle.all():
            break
    return cost, useful


def warp_cost(nA, nB, direct, asteps=2, batch_a=6, bsteps=20, batch_b=8, sort_b=False):
    # task t -> angle t // SPW, sample t % SPW (as in the kernel)
    A = nA.T.reshape(-1)
    B = np.where(direct, 0, nB).T.reshape(-1)
    ca, ua = replay(A, asteps, batch_a, C_A_STEP, C_HAND_A)
    order = np.argsort(-B, kind="stable") if sort_b else None
    cb, ub = replay(B, bsteps, batch_b, C_B_STEP, C_HAND_B, order)
    return ca + cb, ua, ub


def main(warps):
    rng = np.random.default_rng(1)
    a = rng.uniform(-0.5, 0.5, warps * SPW)
    b = rng.uniform(-0.5, 0.5, warps * SPW)
    nA, nB, direct = task_steps(a, b)
    print("mean steps per task: A %.2f  B %.2f (handed-over tasks %.1f %%)" % (
        nA.mean(), np.where(direct, 0, nB).mean(), 100 * (~direct).mean()))

    def total(**kw):
        return sum(warp_cost(nA[w * SPW:(w + 1) * SPW], nB[w * SPW:(w + 1) * SPW],
                             direct[w * SPW:(w + 1) * SPW], **kw)[0] for w in range(warps)) / warps
    base = total()
    print("baseline (2 / 6 / 20 / 8): %.0f warp instructions per 64-sample group" % base)
    for bs in (12, 16, 20, 24, 32):
        print("  polynomial steps per round %2d: %+.1f %%" % (bs, 100 * (total(bsteps=bs) / base - 1)))
    for bb in (4, 8, 12, 16):
        print("  stage-B hand-out batch %2d: %+.1f %%" % (bb, 100 * (total(batch_b=bb) / base - 1)))
    for asx in (1, 2, 3):
        print("  exact steps per round %d: %+.1f %%" % (asx, 100 * (total(asteps=asx) / base - 1)))
    print("  stage B handed out longest-first: %+.1f %% (before the cost of sorting)" % (
        100 * (total(sort_b=True) / base - 1)))
    for bs in (8, 12, 16):
        print("  longest-first with %2d steps per round: %+.1f %%" % (bs, 100 * (total(sort_b=True, bsteps=bs) / base - 1)))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 48)
