#!/bin/bash
P="python experiments/coresidency_probe.py"
X="ASW_STACK=persist"
env $X ASW_CARVE=71 $P -1 0 2>&1 | tail -1
env $X ASW_STACK_SHAPE=512112 ASW_CARVE=71 $P -1 0 2>&1 | tail -1
env $X ASW_STACK_SHAPE=512112 ASW_CARVE=100 $P -1 0 2>&1 | tail -1
env $X ASW_STACK_SHAPE=512010 ASW_CARVE=71 $P -1 0 2>&1 | tail -1
env $X ASW_STACK_SHAPE=384012 ASW_CARVE=71 $P -1 0 2>&1 | tail -1
env $X ASW_STACK_SHAPE=384012 ASW_CARVE=44 ASW_STFT_CTAS=1 $P -1 0 2>&1 | tail -1
env $X ASW_STACK_SHAPE=512012 ASW_CARVE=44 ASW_STFT_CTAS=1 $P -1 0 2>&1 | tail -1
env $X ASW_STACK_SHAPE=384012 ASW_CARVE=71 $P 0 -1 2>&1 | tail -1
