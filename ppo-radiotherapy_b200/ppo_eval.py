"""Evaluation of a saved policy on the device-resident vector env — the reference's ppo_eval.py:5-36 with the same
signature and return value, without the per-step device -> host -> device round trip (ppo_eval.py:25-26).

    from ppo_radiotherapy_b200 import RadiotherapyVectorEnv, PPO
    from ppo_radiotherapy_b200.ppo_eval import evaluate
    envs = RadiotherapyVectorEnv(1024, visionless=True, device="cuda")
    returns = evaluate(envs, 64, "saves/20M.model", eval_episodes=3, Model=PPO, device=torch.device("cuda"))
"""
import numpy as np
import torch


def evaluate(envs, feature_dim: int, model_path: str, eval_episodes: int, Model, device=torch.device("cuda"),
             log=print):
    observation_shape = envs.single_observation_space.shape
    action_space = envs.single_action_space.shape
    agent = Model(observation_shape, action_space, feature_dim).to(device)
    agent.load_state_dict(torch.load(model_path, map_location=device, weights_only=True))     # ppo_eval.py:17
    agent.eval()

    obs, _ = envs.reset(options={"backend": "torch"})
    episodic_returns = []
    if log:
        log("Starting evaluation")
    with torch.no_grad():
        while len(episodic_returns) < eval_episodes:
            actions, _, _, _ = agent.get_action_and_value(obs)
            obs, _, terminated, _, infos = envs.step(actions)
            if bool(terminated.any()):                                      # "episode" in infos (ppo_eval.py:27)
                ep = infos["episode"]
                mean_returns = float(np.mean(ep["r"][ep["_r"]]))
                if log:
                    log(f"eval_episode={len(episodic_returns)}, episodic_return={mean_returns}")
                episodic_returns += [mean_returns]
    return episodic_returns
