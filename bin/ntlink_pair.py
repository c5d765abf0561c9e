#!/usr/bin/env python3
"drop-in for bin/ntlink_pair.py of bcgsc/ntLink: python -m ntlink_b200.pair"
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntlink_b200.pair import main
main()
