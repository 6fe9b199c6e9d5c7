"""Debug helper: one tiny aperiodic + periodic evaluation against the oracle (same as __graft_entry__.smoke)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.smoke()
