"""Protocol-level model of the persistent tcgen05 kernels (few-shot-vid2vid_b200/csrc/conv_tc.cu:k_conv_tc_p and
csrc/spade_tc.cu:k_spade_tc_p): the TMA producer, the MMA issuer and the epilogue warps synchronise through full/empty ring
barriers plus acc_full/acc_empty accumulator barriers with mbarrier phase-parity waits.  The model replays the kernels'
wait/arrive sequences (same stage counter, same parity expressions) under random schedules and checks that they neither
deadlock nor let a ring slot or a TMEM accumulator be overwritten before its consumer is done.  (Written before the conv kernel
first ran on hardware; both kernels have since passed the GPU suite.)  Keep it in sync with the kernels when the protocol
changes.  CPU only."""
import random

class MBar:
    def __init__(self, count): self.count=count; self.pending=count; self.phase=0
    def arrive(self):
        self.pending-=1
        assert self.pending>=0
        if self.pending==0: self.phase^=1; self.pending=self.count
    def test(self, parity):   # try_wait.parity: true if the phase with this parity has completed
        return self.phase != parity

def run(num_tiles, num_k, STG, seed):
    rnd=random.Random(seed)
    full=[MBar(1) for _ in range(STG)]; empty=[MBar(1) for _ in range(STG)]
    acc_full=[MBar(1),MBar(1)]; acc_empty=[MBar(4),MBar(4)]
    slot_owner=[None]*STG      # which (tile,kb) data is in the slot
    acc_state=[None,None]      # ('writing',tile) / ('ready',tile) / None
    log=[]
    def producer():
        it=0
        for tile in range(num_tiles):
            for kb in range(num_k):
                s=it%STG; ph=(it//STG)&1
                while not empty[s].test(ph^1): yield
                assert slot_owner[s] is None, ('slot overwritten', s, slot_owner[s])
                slot_owner[s]=(tile,kb)
                full[s].arrive()          # TMA completes (expect_tx arrive + bytes) -> modelled as one arrive
                it+=1; yield
    def mma():
        it=0
        for i in range(num_tiles):
            buf=i&1
            while not acc_empty[buf].test(((i>>1)&1)^1): yield
            assert acc_state[buf] is None, ('accumulator overwritten', buf, acc_state[buf])
            acc_state[buf]=('writing',i)
            for kb in range(num_k):
                s=it%STG; ph=(it//STG)&1
                while not full[s].test(ph): yield
                assert slot_owner[s]==(i,kb), ('wrong data', slot_owner[s], (i,kb))
                slot_owner[s]=None
                empty[s].arrive()         # tcgen05.commit -> empty
                it+=1; yield
            acc_state[buf]=('ready',i)
            acc_full[buf].arrive(); yield
    done_warps=[[0,0] for _ in range(num_tiles)]
    def epi(w):
        for i in range(num_tiles):
            buf=i&1
            while not acc_full[buf].test((i>>1)&1): yield
            assert acc_state[buf]==('ready',i), ('epilogue reads wrong acc', acc_state[buf], i)
            yield
            done_warps[i][0]+=1
            if done_warps[i][0]==4: acc_state[buf]=None
            acc_empty[buf].arrive(); yield
    procs=[producer(),mma()]+[epi(w) for w in range(4)]
    alive=list(range(len(procs)))
    stall=0
    while alive:
        k=rnd.choice(alive)
        try:
            next(procs[k]); stall+=1
        except StopIteration:
            alive.remove(k); stall=0
        if stall>200000: raise RuntimeError('deadlock/livelock', num_tiles,num_k,STG,seed)
    return True



def test_persistent_conv_barrier_protocol_is_deadlock_and_hazard_free():
    for seed in range(200):
        rnd = random.Random(seed)
        assert run(rnd.randint(1, 9), rnd.randint(1, 12), rnd.randint(2, 6), seed)


def run_spade(num_tiles, kblocks_per_map, STG, ngrp, seed):
    """k_spade_tc_p: per tile the ring carries sum(K_i / 32) K blocks (maps back to back, each map accumulating into its own TMEM
    columns of the tile's buffer); the epilogue is one group of four warps alternating between the two buffers (backward, ngrp = 1)
    or two groups of four warps, group g owning buffer g and every second tile (forward, ngrp = 2).  A group arrives on acc_empty
    as soon as its last tcgen05.ld of the tile is done and stores afterwards."""
    rnd = random.Random(seed)
    num_k = sum(kblocks_per_map)
    full = [MBar(1) for _ in range(STG)]; empty = [MBar(1) for _ in range(STG)]
    acc_full = [MBar(1), MBar(1)]; acc_empty = [MBar(4), MBar(4)]
    slot_owner = [None] * STG
    acc_state = [None, None]
    reads = [0] * num_tiles
    def producer():
        it = 0
        for tile in range(num_tiles):
            for i, kbs in enumerate(kblocks_per_map):
                for kb in range(kbs):
                    s = it % STG; ph = (it // STG) & 1
                    while not empty[s].test(ph ^ 1): yield
                    assert slot_owner[s] is None, ('slot overwritten', s, slot_owner[s])
                    slot_owner[s] = (tile, i, kb)
                    full[s].arrive()
                    it += 1; yield
    def mma():
        it = 0
        for ti in range(num_tiles):
            buf = ti & 1
            while not acc_empty[buf].test(((ti >> 1) & 1) ^ 1): yield
            assert acc_state[buf] is None, ('accumulator overwritten', buf, acc_state[buf])
            acc_state[buf] = ('writing', ti)
            for i, kbs in enumerate(kblocks_per_map):
                for kb in range(kbs):
                    s = it % STG; ph = (it // STG) & 1
                    while not full[s].test(ph): yield
                    assert slot_owner[s] == (ti, i, kb), ('wrong data', slot_owner[s], (ti, i, kb))
                    slot_owner[s] = None
                    empty[s].arrive()
                    it += 1; yield
            acc_state[buf] = ('ready', ti)
            acc_full[buf].arrive(); yield
    def epi(grp, w):
        for ti in range(num_tiles):
            buf = ti & 1
            if ngrp == 2 and buf != grp:
                continue
            yield                                  # x loads issued before the wait
            while not acc_full[buf].test((ti >> 1) & 1): yield
            assert acc_state[buf] == ('ready', ti), ('epilogue reads wrong acc', acc_state[buf], ti)
            yield                                  # tcgen05.ld of every map
            reads[ti] += 1
            if reads[ti] == 4: acc_state[buf] = None
            acc_empty[buf].arrive(); yield         # arrive BEFORE the stores
            yield                                  # staging + global stores
    procs = [producer(), mma()] + [epi(g, w) for g in range(ngrp) for w in range(4)]
    alive = list(range(len(procs)))
    stall = 0
    while alive:
        k = rnd.choice(alive)
        try:
            next(procs[k]); stall += 1
        except StopIteration:
            alive.remove(k); stall = 0
        if stall > 400000: raise RuntimeError('deadlock/livelock', num_tiles, kblocks_per_map, STG, ngrp, seed)
    assert all(r == 4 for r in reads), reads          # every tile drained by exactly four warps
    return True


def test_persistent_spade_barrier_protocol_is_deadlock_and_hazard_free():
    for seed in range(200):
        rnd = random.Random(1000 + seed)
        maps = [rnd.randint(1, 2) for _ in range(rnd.randint(1, 3))]
        assert run_spade(rnd.randint(1, 11), maps, rnd.randint(2, 6), rnd.randint(1, 2), seed)
