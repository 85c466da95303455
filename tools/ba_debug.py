import sys, os
sys.path.insert(0, '/root/repo')
os.environ['SE2GPU_BA_DEBUG'] = '1'
from tools import synth
from se2lam_b200.ba import LocalBA
prob = synth.ba_config('C4')
ba = LocalBA.from_problem(prob)
ba.optimize(10); ba.reset()
print('--- second run'); sys.stdout.flush()
ba.optimize(10)
