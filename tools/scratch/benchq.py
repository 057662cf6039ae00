"""One-line summary of a bench.py JSON line: value, e2e, us per Lanczos step, Lanczos share, steps per solve."""
import json,sys
d=json.load(open(sys.argv[1]));print(sys.argv[1], d["value"],d["e2e"]["value"],d["roofline"]["us_per_lanczos_step"],d["roofline"]["share_of_timed_region"],d["config"]["lanczos_steps_per_solve"])
