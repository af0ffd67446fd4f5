"""Randomised whole-program parity stress on CPU: random inputs (lengths below k, FASTA/FASTQ, equal sizes, single-end),
flags, step sizes, engines per GPU, emulated GPU counts, host thread counts, raw-text vs host-parsed steps and seeding,
table budgets (waves); every output file and counter of the
emulated drop-in binary compared with the oracle CLI.  usage: stress_cli_emu.py SEED SECONDS."""
import os, random, shutil, sys, time
from pathlib import Path
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from tests import cli_cases as cc, oracle_lib as ol
EMU = Path(__file__).resolve().parents[2] / 'tests' / 'emu' / 'nk_emu_cli'
rnd = random.Random(int(sys.argv[1])); budget = float(sys.argv[2])
base = Path(f'/tmp/stress_cli_{sys.argv[1]}'); shutil.rmtree(base, ignore_errors=True); base.mkdir()
t0 = time.time(); n = fails = 0; rcs = {}
while time.time() - t0 < budget:
    d = base / f'c{n}'; d.mkdir()
    fasta = rnd.random() < 0.25
    npairs = rnd.choice([200, 600, 1500, 3000])
    f, r = cc.synth(d, 's', npairs, seed=rnd.randrange(1 << 30), read_len=rnd.choice([60, 100, 150, 250]),
                    equal=rnd.random() < 0.3, fasta=fasta)
    if rnd.random() < 0.4:
        f = cc.mutate_lengths(f, d / ('m_1' + f.suffix), seed=rnd.randrange(1000), fastq=not fasta)
        r = cc.mutate_lengths(r, d / ('m_2' + r.suffix), seed=rnd.randrange(1000), fastq=not fasta)
    k = rnd.choice([5, 9, 15, 21, 25, 31]); p = rnd.choice([1, 2, 3, 4, 8, 16]); depth = max(2 * p, rnd.choice([4, 16, 50, 100, 400]))
    args = ['-f', f] + ([] if rnd.random() < 0.25 else ['-r', r])
    if '-r' not in args: args += ['-s']
    args += ['-k', k, '-p', p, '-d', depth, '-m', 1, '-g', rnd.choice([0.5, 0.9, 0.96, 1.0])]
    if rnd.random() < 0.5: args += ['-c']
    if rnd.random() < 0.3: args += ['-P']
    if fasta: args += ['-t', 'fa', '-o', 'fa']
    elif rnd.random() < 0.3: args += ['-o', 'fa']
    env = {'NKB200_STEP_PAIRS': str(rnd.choice([16, 64, 300, 5000])), 'NKB200_ENGINES_PER_GPU': str(rnd.choice([1, 2, 4])),
           'NK_EMU_DEVICES': str(rnd.choice([1, 2, 3])), 'NK_EMU_SEED': str(rnd.randrange(1 << 30)),
           'NKB200_THREADS': str(rnd.choice([1, 2, 5, 8]))}
    env['NKB200_GPUS'] = env['NK_EMU_DEVICES']
    if rnd.random() < 0.3: del env['NKB200_ENGINES_PER_GPU']          # the default engine policy
    if rnd.random() < 0.2: env['NKB200_HOST_PARSE'] = '1'             # round-1 path: records parsed on the host
    if rnd.random() < 0.2: env['NKB200_HOST_SEED'] = '1'
    if rnd.random() < 0.2: env['NKB200_NO_PREFETCH'] = '1'
    if rnd.random() < 0.2: env['NKB200_EAGER_COUNT'] = '1'            # line ends of the ranges counted before the pipelines start
    if rnd.random() < 0.5: env['NKB200_ROLLING_COUNT'] = '1'          # reverse file counted alongside the first steps
    if rnd.random() < 0.5: env['NKB200_HOT_ENTRIES'] = str(rnd.choice([0, 16, 1024, 65536]))   # hot table off / tiny (collisions) / small
    if rnd.random() < 0.3: env['NKB200_TABLE_BUDGET_MB'] = str(rnd.choice([150, 300, 450, 900]))   # waves, parked tables
    try:
        want = cc.run_cli(ol.ORACLE_CLI, args, d / 'oracle', timeout=600)
        got = cc.run_cli(EMU, args, d / 'emu', env=env, timeout=600)
        rcs[want['rc']] = rcs.get(want['rc'], 0) + 1
        if want['rc'] != 0 and got['rc'] == want['rc']:
            if n < 12: print('rc', want['rc'], want['stderr'][-200:].replace('\n', ' | '), flush=True)
        else:
            cc.assert_same(got, want, 'stress')
        shutil.rmtree(d, ignore_errors=True)
    except Exception as e:
        fails += 1
        print('FAIL', [str(a) for a in args], env, repr(e)[:400], flush=True)
        if fails > 4: break
    n += 1
print('cases', n, 'fails', fails, 'rcs', rcs)
