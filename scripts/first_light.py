"""Diagnostics for the first GPU runs: parity numbers and timings per problem."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt
from oracle import oracle

def run(tag, A, k, p=None):
    if p is None:
        p = configs.markov_table(A, k, 2)
    t0 = time.time()
    m = device.DeviceModel(tag, k)
    tb = time.time() - t0
    dp = torch.from_numpy(p).cuda(); out = torch.empty_like(dp)
    m.rhs(dp, out); torch.cuda.synchronize()
    got = out.cpu().numpy()
    t0 = time.time(); want = oracle.compute_dy_dt(tag, k, p, mode=oracle.MERGED); tc = time.time() - t0
    err = abs(got - want).max() / max(abs(want).max(), 1e-300)
    for _ in range(3): m.rhs(dp, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m.rhs(dp, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ph = m.rhs_profile(dp, out)
    print(f'{tag} A={A} k={k}: relerr={err:.2e} build={tb*1e3:.1f}ms {m.timing} rhs={ms*1e3:.1f}us phases_ms={ph} cpu_merged={tc*1e3:.2f}ms info={m.info}', flush=True)

if __name__ == '__main__':
    run('__canary_problem_radioactive_decay', 2, 3, numpy.full(8, 0.125))
    for tag, A, k in [('ex2-ferromagnetic-chain', 2, 7), ('ex3-copolymerization', 4, 6), ('ex5-msrtf-machine', 5, 5),
                      ('ex4-chemical-turing', 9, 5), ('ex4var2-chemical-turing', 10, 5)]:
        run(tag, A, k)
    for A, k, R in [(10, 5, 8), (10, 6, 8), (10, 7, 8)]:
        rules = configs.random_rule_set(A, R, seed=1)
        tag = configs.synthetic_tag(A, R, 1)
        oracle.register_rules(tag, A, rules); mt.register_rule_set(tag, A, rules)
        run(tag, A, k, configs.dirichlet_product_table(A, k, 3))
