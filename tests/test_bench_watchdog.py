"""bench.py's progress watchdog (host logic, no GPU): when a block after the headline stops completing timed regions, rank 0 still
emits the JSON line it has, with the stalled block named, and the process leaves with status 0; a stall before the headline leaves
with status 3; a run that keeps beating is left alone."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = textwrap.dedent('''
    import json, os, sys, time
    sys.path.insert(0, {root!r})
    import bench
    bench.Watchdog.LIMITS = (0.6, 1.5)        # (optional block, mandatory block) seconds, instead of 240 / 600
    bench.Watchdog.POLL = 0.05
    r, w = os.pipe()
    wd = bench.Watchdog(); wd.fd = w; wd.rank = 0
    mode = sys.argv[1]
    if mode == "optional":
        wd.finalize = lambda aborted=None: {{"value": 1.0, "aborted": aborted}}
        wd.enter("large4k", True)
        time.sleep(10)
    elif mode == "mandatory":
        wd.enter("headline", False)
        time.sleep(10)
    elif mode == "beating":
        wd.finalize = lambda aborted=None: {{"value": 1.0, "aborted": aborted}}
        wd.enter("large4k", True)
        for _ in range(30):
            time.sleep(0.1); wd.beat()
        print("finished")
''')


def _run(mode):
    env = dict(os.environ, PYTHONPATH=ROOT)
    code = PROG.format(root=ROOT)
    # the watchdog writes the line to the fd it was given (a pipe here); read it back through /proc is not possible after exit, so the
    # child dups the pipe's write end onto its stdout
    code = code.replace("r, w = os.pipe()", "r, w = None, 1")
    return subprocess.run([sys.executable, "-c", code, mode], capture_output=True, text=True, timeout=60, env=env)


def test_stall_in_optional_block_still_prints_the_line():
    p = _run("optional")
    assert p.returncode == 0, p.stderr
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.0 and "large4k" in line["aborted"]
    assert "bench watchdog" in p.stderr


def test_stall_before_the_headline_exits_nonzero():
    p = _run("mandatory")
    assert p.returncode == 3 and p.stdout.strip() == ""


def test_progress_keeps_the_watchdog_quiet():
    p = _run("beating")
    assert p.returncode == 0 and p.stdout.strip() == "finished" and "watchdog" not in p.stderr
