"""Randomised engine-level parity stress on CPU: random k / depth / capacity / partitions / read shapes / list sizes /
launch orders through the emulation backend (tests/emu), every step compared with the oracle slot for slot
(tests/engine_cases.run_case).  usage: stress_engine_emu.py SEED SECONDS.  Found the "step overfills the table" case."""
import ctypes, os, random, sys, time, traceback
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from nomalise_kmers_multi_large_b200 import capi
from tests import engine_cases as ec
lib = ctypes.CDLL(str(__import__('pathlib').Path(__file__).resolve().parents[2] / 'tests' / 'emu' / 'libnk_emu.so')); capi._declare_engine(lib); capi._declare_pipeline(lib)
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
t0 = time.time(); n = 0; fails = 0
while time.time() - t0 < float(sys.argv[2] if len(sys.argv) > 2 else 600):
    cfg = ec.random_case(rnd)
    os.environ['NK_EMU_SEED'] = str(rnd.randrange(1 << 30))
    if rnd.random() < 0.2:
        os.environ['NKB200_OPEN_FRAC'] = '0.05'; os.environ['NKB200_PEND_FRAC'] = '0.1'
    else:
        os.environ.pop('NKB200_OPEN_FRAC', None); os.environ.pop('NKB200_PEND_FRAC', None)
    os.environ['NKB200_HOT_ENTRIES'] = str(rnd.choice([0, 16, 256, 4096, 1 << 20]))   # hot table off / colliding / roomy
    if os.environ.get('NK_STRESS_VERBOSE'):
        print('case', cfg, os.environ.get('NK_EMU_SEED'), os.environ.get('NKB200_OPEN_FRAC'), flush=True)
    try:
        ec.run_case(lib, **cfg)
    except Exception as e:
        fails += 1
        print('FAIL', cfg, os.environ.get('NK_EMU_SEED'), os.environ.get('NKB200_OPEN_FRAC'), repr(e)[:300], flush=True)
        if fails > 5: break
    n += 1
print('cases', n, 'fails', fails)
