import heapq, random, sys
import numpy as np
def sim(G=296, n_tiles=13000, sigma=0.15, slow_frac=0.0, slow_factor=1.3, defer=1.0, early_ticket=True, flush=0.15, seed=1):
    """Each CTA: loop { (ticket for NEXT taken at start of current tile if early_ticket else at its end);
    work ~ lognormal(mean 1, sigma) * speed[c]; publish; flush of the tile `defer` tiles back needs all earlier tiles published }.
    defer=0: flush right after publish; defer=1: after the next tile's work; defer=0.3: after 30% of next tile's work."""
    rng = np.random.default_rng(seed)
    speed = np.ones(G); 
    ns = int(G*slow_frac); speed[:ns] = slow_factor
    rng.shuffle(speed)
    pub = np.full(n_tiles, np.inf)
    next_ticket = 0
    # state per CTA: time, current tile, next tile
    t = np.zeros(G); cur = np.full(G, -1); nxt = np.full(G, -1)
    # event-driven: process CTAs in order of time
    heap = [(0.0, c, 0) for c in range(G)]  # (time, cta, phase) phase 0 = start (take first ticket)
    heapq.heapify(heap)
    pending = {}  # cta -> (tile, ready_time) waiting flush
    done_t = 0.0
    pubmax_prefix = []  # we need max publish time over tiles < i : maintain lazily
    # simpler: iterate in rounds since dependencies need publish times of lower tiles, which are produced by events earlier in time mostly.
    # We'll process with a loop that may need to wait: implement as generator-like state machine.
    state = [dict(tile=-1, nxt=-1, pend=None) for _ in range(G)]
    def take():
        nonlocal next_ticket
        v = next_ticket; next_ticket += 1; return v
    # initial
    events = []
    for c in range(G):
        state[c]['tile'] = take()
    for c in range(G):
        if early_ticket: state[c]['nxt'] = take()
        heapq.heappush(events, (0.0, c, 'work'))
    published = np.full(n_tiles + 4*G, np.inf)
    prefix_ok_upto = 0  # all tiles < prefix_ok_upto published; prefix_time = max publish time among them
    prefix_time = 0.0
    waiting = []  # (needed_tile_index, cta, time_ready, kind)
    finish = 0.0
    def all_pub_before(i):
        nonlocal prefix_ok_upto, prefix_time
        while prefix_ok_upto < i and published[prefix_ok_upto] < np.inf:
            prefix_time = max(prefix_time, published[prefix_ok_upto]); prefix_ok_upto += 1
        return prefix_ok_upto >= i
    tiles_done = 0
    while events or waiting:
        if not events:
            raise RuntimeError("deadlock")
        tm, c, kind = heapq.heappop(events)
        st = state[c]
        if kind == 'work':
            tile = st['tile']
            if tile >= n_tiles:
                # flush pending then exit
                if st['pend'] is not None:
                    waiting.append((st['pend'], c, tm, 'final'))
                    st['pend'] = None
                else:
                    finish = max(finish, tm)
                # process waiting list below
            else:
                dur = speed[c] * rng.lognormal(-0.5*sigma**2, sigma)
                if defer > 0 and defer < 1 and st['pend'] is not None:
                    # partial deferral: do defer*dur of work, then flush pending, then rest
                    waiting.append((st['pend'], c, tm + defer*dur, ('partial', tile, (1-defer)*dur)))
                    st['pend'] = None
                else:
                    heapq.heappush(events, (tm + dur, c, 'published'))
        elif kind == 'published':
            tile = st['tile']
            published[tile] = tm
            # next ticket
            if early_ticket:
                st['tile'] = st['nxt']; st['nxt'] = take()
            else:
                st['tile'] = take()
            if defer == 0:
                waiting.append((tile, c, tm, 'now'))
            elif defer >= 1:
                if st['pend'] is not None:
                    waiting.append((st['pend'], c, tm, ('then_store', tile)))
                else:
                    st['pend'] = tile
                    heapq.heappush(events, (tm, c, 'work'))
            else:
                st['pend'] = tile
                heapq.heappush(events, (tm, c, 'work'))
        # try to release waiters
        still = []
        for (need, cc, tr, k) in waiting:
            if all_pub_before(need):
                # compute time when all earlier published: max over published[:need]
                tmax = published[:need].max() if need > 0 else 0.0
                tgo = max(tr, tmax) + flush
                tiles_done += 1
                if k == 'now':
                    heapq.heappush(events, (tgo, cc, 'work'))
                elif k == 'final':
                    finish = max(finish, tgo)
                elif k[0] == 'then_store':
                    state[cc]['pend'] = k[1]
                    heapq.heappush(events, (tgo, cc, 'work'))
                elif k[0] == 'partial':
                    heapq.heappush(events, (tgo + k[2], cc, 'published'))
            else:
                still.append((need, cc, tr, k))
        waiting = still
    ideal = n_tiles * (1 + flush) * speed.mean() / G   # not exactly ideal with dynamic scheduling
    return finish, ideal
for sigma in (0.1, 0.25):
  for slow in ((0.0,1.0),(0.1,1.3)):
    for defer in (0, 0.3, 1.0):
      for early in (True, False):
        f, ideal = sim(sigma=sigma, slow_frac=slow[0], slow_factor=slow[1], defer=defer, early_ticket=early, n_tiles=6000)
        print(f"sigma {sigma} slow {slow} defer {defer} early {early}: time {f:.2f} ideal {ideal:.2f} eff {ideal/f:.2f}")
