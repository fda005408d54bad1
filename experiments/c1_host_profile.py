import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.mic_array import Mic_Array
scene = synth.desk_array(7, np.random.default_rng(0), 44100)
ma = Mic_Array(scene.mic_positions, Spk_Range=scene.roi, fs=44100)
mix = torch.from_numpy(synth.mixture(scene, 3, 132300, 0))
for _ in range(5): ma.Apply_SRP_PHAT(mix)
torch.cuda.synchronize()
ts=[]
for i in range(50):
    t=time.perf_counter(); p,_=ma.Apply_SRP_PHAT(mix); torch.cuda.synchronize(); ts.append(time.perf_counter()-t)
print('median us', np.median(ts)*1e6, 'min', min(ts)*1e6, 'patches', len(p))
pr=cProfile.Profile(); pr.enable()
for i in range(200): ma.Apply_SRP_PHAT(mix)
torch.cuda.synchronize(); pr.disable()
st=pstats.Stats(pr); st.sort_stats('cumulative').print_stats(28)
