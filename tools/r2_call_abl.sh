#!/bin/bash
timeout 300 python tools/cpu_profile_step.py 2>&1 | grep -v Warning | head -75 | cut -c1-190
