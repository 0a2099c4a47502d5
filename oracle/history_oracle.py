"""oracle/history_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

pandas restatement of the history join + resampling of the reference predictor wrapper
(sicnav_diffusion/JMID/mid_sim_wrapper.py:244-298: `_gen_agent_df` up to the sub-sampled frame table, `subsample_df`), i.e. the
same pandas calls in the same order on the same column names.  pandas is a dependency of the reference itself and is present
in this image; the checker of snb/jmid/history.py.

Only tests/ may import this module.
"""
import numpy as np
import pandas as pd


def subsample_df(scene_df, time_step):
    """mid_sim_wrapper.py:283-298."""
    subsampled_df = scene_df
    subsampled_df["datetime"] = pd.to_datetime(subsampled_df["time"] * 100)
    subsampled_df = subsampled_df.resample(f"{int(round(time_step * 100))}ns", on="datetime", origin="end").last()
    subsampled_df = subsampled_df.interpolate(method="linear", axis=0)
    return subsampled_df


def gen_agent_frames(prev_states, prev_robot_states, time_step, num_hist_frames):
    """mid_sim_wrapper.py:253-272 -> (humans [H, F, 2], robot [F, 2]) float64, F <= num_hist_frames frames, oldest first
    (the rows expand_df (:301-310) walks)."""
    num_hums = len(prev_states)
    individual_dfs = []
    for hum_idx in range(num_hums):
        individual_dfs.append(pd.DataFrame(prev_states[hum_idx], columns=["pos_x_" + str(hum_idx), "pos_y_" + str(hum_idx), "time"]))
    robot_df = pd.DataFrame(prev_robot_states, columns=["pos_x_robot", "pos_y_robot", "time"])
    agent_df = individual_dfs[0].set_index("time")
    for hum_idx in range(1, num_hums):
        agent_df = agent_df.join(individual_dfs[hum_idx].set_index("time"))
    agent_df = agent_df.join(robot_df.set_index("time"))
    agent_df = agent_df.dropna()
    agent_df = agent_df.reset_index()
    agent_df = agent_df.sort_values(by=["time"])
    sub = subsample_df(agent_df, time_step).tail(num_hist_frames)
    hum = np.stack([np.stack([sub["pos_x_" + str(i)].to_numpy(), sub["pos_y_" + str(i)].to_numpy()], 1) for i in range(num_hums)], 0)
    rob = np.stack([sub["pos_x_robot"].to_numpy(), sub["pos_y_robot"].to_numpy()], 1)
    return hum.astype(np.float64), rob.astype(np.float64)
