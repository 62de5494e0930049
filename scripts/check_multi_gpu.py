"""Run under torchrun on N GPUs: the sharded, overlapped right-hand side must equal the unsplit one."""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch, torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt, parallel
import bench

A, k, R = 10, 6, 4 * world
rules = configs.random_rule_set(A, R, seed=5)
mt.register_rule_set('mg-full', A, rules)
mt.register_rule_set(f'mg-part{rank}', A, parallel.split_rule_set(rules, world, rank))
full = device.DeviceModel('mg-full', k)
part = device.DeviceModel(f'mg-part{rank}', k)
n = A ** k
p = bench.device_product_table(A, k, 9, dev)
want = full.rhs(p)
for chunks in (1, 3, 8):
    sh = parallel.OverlappedRhs(part.weights, part.flux_rows, n, chunks=chunks, device=dev)
    pf = torch.zeros(sh.padded, dtype=torch.float64, device=dev); pf[:n] = p
    out = torch.zeros_like(pf)
    for _ in range(3):
        sh.rhs_full(pf, out)
    torch.cuda.synchronize()
    err = float((out[:n] - want).abs().max() / want.abs().max())
    print(f'rank {rank}/{world} chunks={chunks} max rel dev vs unsplit = {err:.2e}', flush=True)
    assert err < 1e-13
for chunks in (1, 5):
    ar = parallel.OverlappedAllReduceRhs(part.weights, part.flux_rows, n, chunks=chunks)
    out = torch.zeros(n, dtype=torch.float64, device=dev)
    for _ in range(2):
        ar.rhs_full(p, out)
    torch.cuda.synchronize()
    err = float((out - want).abs().max() / want.abs().max())
    print(f'rank {rank}/{world} all-reduce blocks={chunks} max rel dev vs unsplit = {err:.2e}', flush=True)
    assert err < 1e-13
for rounds in (1, 4):
    peer = parallel.PeerExchangeRhs(part, rounds=rounds)
    for _ in range(3):
        got = peer.rhs_full(p)
    peer.check()
    err = float((got[:n] - want).abs().max() / want.abs().max())
    everyone = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(everyone, got[:n].contiguous())
    same = all(bool(torch.equal(everyone[0], e)) for e in everyone)
    print(f'rank {rank}/{world} peer exchange rounds={rounds} max rel dev vs unsplit = {err:.2e}, '
          f'identical on all ranks: {same}', flush=True)
    assert err < 1e-13 and same
    if rounds == 4:
        # the stepper of all ranks together against the stepper of one rank on the whole rule set
        p0 = p.cpu().numpy()
        ts = numpy.linspace(0.0, 2.0, 5)
        seqs = [[1], [2, 3], [0, 0, 1]]
        kw = dict(size_a=A, cl_k=k, p0=p0, ts=ts, rtol=1e-9, atol=1e-12, observables=seqs)
        together, st_t = mt.ode_integrate_device(tag=f'mg-part{rank}', peer_group=peer, want_stats=True, **kw)
        alone, st_a = mt.ode_integrate_device(tag='mg-full', want_stats=True, **kw)
        dev_states = float(abs(together[0] - alone[0]).max() / abs(alone[0]).max())
        print(f'rank {rank}/{world} peer stepper vs single-rank stepper: states {dev_states:.2e}, '
              f'observables {float(abs(together[1] - alone[1]).max()):.2e}, steps {st_t} vs {st_a}', flush=True)
        assert dev_states < 1e-11 and st_t['accepted'] == st_a['accepted']
    peer.close()
# a problem of the reference's own set dealt to the ranks by the library (tapes_model_part): the
# program's leaf worlds, not rewrite rules, are the unit
for b_tag, b_a, b_k in (('ex4-chemical-turing', 9, 5), ('ex3-copolymerization', 4, 9)):
    whole = device.DeviceModel(b_tag, b_k)
    share = device.DeviceModel(b_tag, b_k, part=(rank, world))
    bn = b_a ** b_k
    bp = torch.from_numpy(configs.markov_table(b_a, b_k, 31)).to(dev)
    b_want = whole.rhs(bp)
    peer = parallel.PeerExchangeRhs(share, rounds=2)
    for _ in range(2):
        got = peer.rhs_full(bp)
    peer.check()
    err = float((got[:bn] - b_want).abs().max() / b_want.abs().max())
    rules_here = torch.tensor([share.info['n_flux_rules']], device=dev)
    dist.all_reduce(rules_here)
    print(f'rank {rank}/{world} {b_tag} k={b_k}: {share.info["n_flux_rules"]} of {whole.info["n_flux_rules"]} flux rules here, '
          f'peer exchange max rel dev vs one GPU = {err:.2e}', flush=True)
    assert err < 1e-13 and int(rules_here.item()) == whole.info['n_flux_rules']
    if b_tag.startswith('ex4'):
        kw = dict(tag=b_tag, size_a=b_a, cl_k=b_k, p0=configs.markov_table(b_a, b_k, 31), ts=numpy.linspace(0.0, 5.0, 3),
                  rtol=1e-9, atol=1e-12, observables=[[1], [2, 3]])
        together = mt.ode_integrate_device(peer_group=peer, **kw)
        alone = mt.ode_integrate_device(**kw)
        dev_states = float(abs(together[0] - alone[0]).max() / abs(alone[0]).max())
        print(f'rank {rank}/{world} {b_tag}: stepper of all ranks vs one rank: states {dev_states:.2e}', flush=True)
        assert dev_states < 1e-11
    peer.close()
plain = parallel.ShardedRhs(lambda a, b: part.rhs(a, b), n, device=dev)
pf = torch.zeros(plain.padded, dtype=torch.float64, device=dev); pf[:n] = p
out = torch.zeros_like(pf)
plain.rhs_full(pf, out)
torch.cuda.synchronize()
assert float((out[:n] - want).abs().max() / want.abs().max()) < 1e-13
dist.destroy_process_group()
if rank == 0:
    print('multi-gpu check ok')
