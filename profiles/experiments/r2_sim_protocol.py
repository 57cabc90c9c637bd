"""Discrete-event simulation of sim_small_kernel's mbarrier protocol (parity waits!) with random delays."""
import heapq, random, sys
S = int(sys.argv[1]) if len(sys.argv) > 1 else 5
EARLY = (sys.argv[2] == "early") if len(sys.argv) > 2 else True
R, G, TEAMS = 6, 4, 6
NT = 60                      # tiles in this CTA
SEG = [0, 17, 40, NT]        # segment boundaries (image ends) in local tiles
random.seed(int(sys.argv[3]) if len(sys.argv) > 3 else 0)

class MBar:
    def __init__(s, count): s.count = count; s.pending = count; s.phase = 0; s.completions = 0
    def arrive(s):
        s.pending -= 1
        assert s.pending >= 0, "over-arrival"
        if s.pending == 0: s.pending = s.count; s.phase ^= 1; s.completions += 1
    def test(s, parity): return s.phase != parity

ring_full = [MBar(1) for _ in range(R)]; ring_empty = [MBar(G) for _ in range(R)]
k_full = [MBar(16) for _ in range(S)]; k_empty = [MBar(1) for _ in range(S)]
s_full = [MBar(1), MBar(1)]; p_full = [MBar(1), MBar(1)]; o_done = [MBar(1), MBar(1)]; o_free = MBar(4)
ring_content = [None] * R          # group id held
stage_rows = [[None] * 16 for _ in range(S)]   # (tile,row) written
stage_inuse = [None] * S           # tile currently being read by MMA
errors = []
now = 0.0
events = []
def at(t, fn): heapq.heappush(events, (t, random.random(), fn))
def delay(a, b): return random.uniform(a, b)

def proc(gen):
    def step():
        try:
            w = next(gen)
        except StopIteration:
            return
        if w[0] == "sleep": at(now + w[1], step)
        else:
            bar, parity = w[1], w[2]
            def poll():
                if bar.test(parity): step()
                else: at(now + 0.05, poll)
            poll()
    step()

def producer():
    n_groups = NT * 4
    for rg in range(n_groups):
        slot = rg % R
        yield ("wait", ring_empty[slot], ((rg // R) & 1) ^ 1)
        # TMA lands later
        def land(slot=slot, rg=rg):
            ring_content[slot] = rg
            ring_full[slot].arrive()
        at(now + delay(0.5, 1.5), land)
        yield ("sleep", 0.02)

def converter(team, w):
    n_groups = NT * 4
    for rg in range(team, n_groups, TEAMS):
        lt, q4 = divmod(rg, 4)
        st = lt % S
        yield ("wait", k_empty[st], ((lt // S) & 1) ^ 1)
        slot = rg % R
        yield ("wait", ring_full[slot], (rg // R) & 1)
        if ring_content[slot] != rg: errors.append(("ring content", now, rg, ring_content[slot], team, w))
        if EARLY:
            ring_empty[slot].arrive()
        yield ("sleep", delay(0.3, 3.0))
        if stage_inuse[st] is not None and stage_inuse[st] != lt: errors.append(("WAR stage", now, lt, stage_inuse[st]))
        stage_rows[st][q4 * 4 + w] = lt
        k_full[st].arrive()
        if not EARLY:
            ring_empty[slot].arrive()

def mma():
    def issue_s(lt):
        st = lt % S
        yield ("wait", k_full[st], (lt // S) & 1)
        bad = [r for r in range(16) if stage_rows[st][r] != lt]
        if bad: errors.append(("RAW S", now, lt, bad, list(stage_rows[st])))
        stage_inuse[st] = lt
        def done(lt=lt): s_full[lt & 1].arrive()
        at(now + 0.1, done)
    lt0 = 0
    for seg in range(len(SEG) - 1):
        nt = SEG[seg + 1] - SEG[seg]
        if seg > 0: yield ("wait", o_free, (seg - 1) & 1)
        if lt0 == 0: yield from issue_s(0)
        for i in range(nt):
            lt = lt0 + i
            if lt + 1 < NT: yield from issue_s(lt + 1)
            st, sb = lt % S, lt & 1
            yield ("wait", p_full[sb], (lt >> 1) & 1)
            bad = [r for r in range(16) if stage_rows[st][r] != lt]
            if bad: errors.append(("RAW O", now, lt, bad))
            def done(st=st, sb=sb, lt=lt):
                stage_inuse[st] = None
                k_empty[st].arrive(); o_done[sb].arrive()
            at(now + 0.15, done)
            yield ("sleep", 0.02)
        lt0 += nt

def softmax():
    lt0 = 0
    for seg in range(len(SEG) - 1):
        nt = SEG[seg + 1] - SEG[seg]
        for i in range(nt):
            lt = lt0 + i
            yield ("wait", s_full[lt & 1], (lt >> 1) & 1)
            yield ("sleep", delay(0.1, 0.3))
            if lt >= 2: yield ("wait", o_done[lt & 1], ((lt - 2) >> 1) & 1)
            p_full[lt & 1].arrive()
        last = lt0 + nt - 1
        yield ("wait", o_done[last & 1], (last >> 1) & 1)
        yield ("sleep", delay(2.0, 6.0))           # epilogue readout
        for _ in range(4): o_free.arrive()
        lt0 += nt

procs = [producer(), mma(), softmax()] + [converter(t, w) for t in range(TEAMS) for w in range(G)]
for g in procs: proc(g)
while events:
    t, _, fn = heapq.heappop(events)
    now = t
    fn()
    if now > 5000: break
print("S", S, "early" if EARLY else "late", "errors", len(errors), "time", round(now, 1))
for e in errors[:6]: print("  ", e)
print("completions k_full", [b.completions for b in k_full], "tiles", NT)
