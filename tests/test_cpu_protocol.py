"""Protocol-level model of the persistent tcgen05 conv kernel (few-shot-vid2vid_b200/csrc/conv_tc.cu:k_conv_tc_p): the TMA
producer, the MMA issuer and the four epilogue warps synchronise through full/empty ring barriers plus
acc_full/acc_empty accumulator barriers with mbarrier phase-parity waits.  The kernel has not run on hardware yet; this
model replays its wait/arrive sequence (same stage counter, same parity expressions) under random schedules and checks
that it neither deadlocks nor lets a ring slot or a TMEM accumulator be overwritten before its consumer is done.  Keep it
in sync with the kernel when the protocol changes.  CPU only."""
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
