import sys; sys.path.insert(0, "/root/repo")
from open_spiel_coup_b200.spiel import load_game
state = load_game("coup").new_initial_state()
from open_spiel_coup_b200.vector_env import CoupVectorEnv
env = CoupVectorEnv(1 << 16, seed=1234, auto_reset=True)
env.step(env.sample_uniform()); x = env.information_state_tensor(); print(x.shape)
from open_spiel_coup_b200.deep_cfr import DeepCFRSolver
policy_net, adv_losses, policy_loss = DeepCFRSolver(num_iterations=2, num_traversals=300, sampling_method="outcome",
                                                    advantage_network_layers=(64, 64), policy_network_layers=(64, 64)).solve()
print("deep cfr", adv_losses[0], policy_loss)
from open_spiel_coup_b200 import agents, mccfr
nfsp = agents.train_nfsp(2000, [64, 64], num_envs=512, reservoir_buffer_capacity=100_000, anticipatory_param=0.1)
print("nfsp", [a.loss for a in nfsp])
solver = mccfr.OutcomeSamplingSolver(num_envs=512); solver.iteration()
print(agents.rl_resp(exploitee=solver.average_policy(), num_train_episodes=1024, eval_every=512, eval_episodes=256, num_envs=256)[-1])
