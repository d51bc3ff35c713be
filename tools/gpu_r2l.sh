#!/bin/bash
timeout 600 python tools/lockstep_profile.py 4 2>&1 | tail -12
