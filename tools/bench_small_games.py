import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from knightvision_b200.selfplay import SelfPlay, engine_for
from knightvision_b200.model import ChessNet
torch.manual_seed(0)
net = ChessNet().eval()
for K in (1, None):
    sp = SelfPlay(net, 5, torch.device("cuda:0"), sims=800, max_plies=4, inflight=K)
    sp.play()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = sp.play(game_id_base=100)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("games 5, sims 800, 4 plies, inflight", sp.inflight, "-> %.3f s" % dt, "waves", sp.eng.mcts_waves(), st["plies"], flush=True)
