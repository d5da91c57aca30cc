"""experiment_0: the reference's rollout driver (experiment/experiment_0.py:11-59) for a batch of
environments, and its data product.

The reference runs 32 threads, each stepping its own Scene 1e5 times with random actions, collecting
per-step records [j1, j2, a0, a1, step_reward, done, collision] into episodes (a list is closed and the scene
reset whenever done or collision), and writes every completed episode as one `"%s\\n" % episode` line of
`data_list.txt` (experiment_0.py:20-34,54-57).  Here env e plays the role of thread e:

  * `run_experiment`        -- the fused-rollout form: K-step kernel launches, records streamed to pinned host
                               memory (joints as float32, the kernel's record format);
  * `run_experiment_exact`  -- one `BatchedScene.step` + masked `reset` per step, float64 joints: what a small
                               run needs to reproduce the reference's text dump digit for digit;
  * `Trajectories`          -- the step records as structure-of-arrays [K, N] numpy arrays, episode segmentation,
                               a binary file format (np.savez: arrays + episode index) and the exporter to the
                               reference's `data_list.txt` text format.

`Trajectories` is host-side numpy (no device needed); everything that computes a step is CUDA behind the C ABI.
"""
import numpy as np

FLAG_COLLISION, FLAG_DONE = 1, 2


class Trajectories:
    """Step records of N environments over K steps, post-step / pre-reset like the reference's record.append.

    j1, j2, a0, a1, reward: [K, N] float arrays; flags: [K, N] uint8 (bit0 collision, bit1 done).
    An episode of env e is a maximal run of steps ending at a step with flags != 0; steps after the last terminal
    step belong to an episode that has not finished and are not exported (experiment_0.py:30-32 appends on
    termination only)."""

    FIELDS = ("j1", "j2", "a0", "a1", "reward", "flags")

    def __init__(self, j1, j2, a0, a1, reward, flags):
        self.j1, self.j2, self.a0, self.a1 = (np.asarray(x) for x in (j1, j2, a0, a1))
        self.reward = np.asarray(reward)
        self.flags = np.asarray(flags, dtype=np.uint8)
        shapes = {np.shape(getattr(self, f)) for f in self.FIELDS}
        if len(shapes) != 1 or len(next(iter(shapes))) != 2:
            raise ValueError("all fields must share one [K, N] shape, got %s" % shapes)
        self.K, self.N = self.flags.shape

    # ---- construction -------------------------------------------------------------------------
    @classmethod
    def from_rollout(cls, rec, actions):
        """rec: the dict BatchedScene.rollout / rollout_host fills ([K,N] tensors, device or pinned host);
        actions: the [K,N,2] tensor/array that was fed to it."""
        def host(x):
            return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
        a = host(actions)
        fl = host(rec["flags"])
        if rec.get("reward") is not None:
            rw = host(rec["reward"])
        else:                                       # compact records: reward is a function of the flags
            rw = np.where(fl & FLAG_DONE, np.float32(1e4), np.where(fl & FLAG_COLLISION, np.float32(-1e3), np.float32(0)))
        return cls(host(rec["j1"]), host(rec["j2"]), a[:, :, 0], a[:, :, 1], rw, fl)

    @classmethod
    def concatenate(cls, parts):
        """chunks of consecutive steps of the same environments -> one Trajectories"""
        return cls(*[np.concatenate([getattr(p, f) for p in parts], axis=0) for f in cls.FIELDS])

    # ---- episodes -----------------------------------------------------------------------------
    def episode_index(self):
        """Completed episodes as an int64 array [E, 3] of (env, first step, last step), ordered by env, then time
        (the order of experiment_0.py:55-57: thread by thread, episode by episode)."""
        t_idx, e_idx = np.nonzero(self.flags.T != 0)[::-1]      # terminal (env, step) pairs sorted by env, then step
        if e_idx.size == 0:
            return np.zeros((0, 3), dtype=np.int64)
        first = np.empty_like(t_idx)
        new_env = np.ones(e_idx.size, dtype=bool)
        new_env[1:] = e_idx[1:] != e_idx[:-1]
        first[new_env] = 0
        first[~new_env] = t_idx[:-1][~new_env[1:]] + 1
        return np.stack([e_idx, first, t_idx], axis=1).astype(np.int64)

    def episodes(self, env=None):
        """iterate (env, first, last) over completed episodes (of one env if given)"""
        idx = self.episode_index()
        if env is not None:
            idx = idx[idx[:, 0] == env]
        for e, a, b in idx:
            yield int(e), int(a), int(b)

    def stats(self):
        idx = self.episode_index()
        term = self.flags[idx[:, 2], idx[:, 0]] if len(idx) else np.zeros(0, dtype=np.uint8)
        return dict(episodes=int(len(idx)), collisions=int(((term & FLAG_COLLISION) != 0).sum()),
                    successes=int(((term & FLAG_DONE) != 0).sum()),
                    ep_len_sum=int((idx[:, 2] - idx[:, 1] + 1).sum()) if len(idx) else 0)

    # ---- binary format ------------------------------------------------------------------------
    def save(self, path):
        """structure-of-arrays binary dump (np.savez): the six [K,N] arrays + the episode index"""
        np.savez(path, episode_index=self.episode_index(), **{f: getattr(self, f) for f in self.FIELDS})

    @classmethod
    def load(cls, path):
        with np.load(path) as z:
            return cls(*[z[f] for f in cls.FIELDS])

    # ---- the reference's text format ------------------------------------------------------------
    @staticmethod
    def _num(x, numpy2):
        r = repr(float(x))
        return "np.float64(%s)" % r if numpy2 else r

    def format_episode(self, env, first, last, numpy2=True):
        """One line of data_list.txt: `"%s" % episode` with episode = [[j1, j2, a0, a1, reward, done, collision], ...]
        (experiment_0.py:23-25).  j1, j2, a0, a1 are numpy float64 scalars in the reference (repr `np.float64(x)` under
        numpy >= 2, `x` before); step_reward is the int 0 until a terminal step sets the float -1000.0 / 10000.0
        (scene_0.py:40,96,99); done / collision are Python bools."""
        recs = []
        for t in range(first, last + 1):
            fl = int(self.flags[t, env])
            rw = float(self.reward[t, env])
            rws = "0" if (rw == 0.0 and fl == 0) else repr(rw)
            recs.append("[%s, %s, %s, %s, %s, %s, %s]" % (
                self._num(self.j1[t, env], numpy2), self._num(self.j2[t, env], numpy2),
                self._num(self.a0[t, env], numpy2), self._num(self.a1[t, env], numpy2), rws,
                "True" if fl & FLAG_DONE else "False", "True" if fl & FLAG_COLLISION else "False"))
        return "[" + ", ".join(recs) + "]"

    def export_text(self, path, numpy2=None):
        """Write data_list.txt (experiment_0.py:54-57): env by env, one completed episode per line."""
        if numpy2 is None:
            numpy2 = int(np.__version__.split(".")[0]) >= 2
        n = 0
        with open(path, "w") as f:
            for e, a, b in self.episodes():
                f.write(self.format_episode(e, a, b, numpy2) + "\n")
                n += 1
        return n


def run_experiment(scene, steps, chunk_steps=64, actions=None, pipeline_steps=4):
    """experiment_0.py:20-34 for every env of `scene` with the fused rollout kernel, `chunk_steps` steps per call, the
    records streamed to pinned host memory by the pipelined host path (copies overlap the kernels):

      * actions=None  -- the actions are drawn in the kernel (Philox stream 0, float64 (u-0.5)*0.1); nothing goes host ->
        device, the joint planes stream out and reward / flags arrive as an event list (`rollout_events_host`); the
        recorded actions a0 / a1 are reproduced on the host from the draw counters (`philox_actions`);
      * actions [steps,N,2] float32 -- host actions in, the four record planes out (`rollout_host`).

    Returns Trajectories (joints as float32) -- the trajectory sink of the fast path."""
    import torch
    parts, t = [], 0
    while t < steps:
        k = min(chunk_steps, steps - t)
        if actions is None:
            sc0 = scene.step_ctr.cpu().numpy().view(np.uint32).astype(np.uint64)
            sink = scene.alloc_event_sink(k, pinned_host=True)
            scene.rollout_events_host(k, sink, chunk_steps=pipeline_steps)
            if sink["count"] > sink["events"].shape[0]:
                raise RuntimeError("event sink overflow: %d events, capacity %d" % (sink["count"], sink["events"].shape[0]))
            reward, flags = scene.events_to_planes(sink, k)
            act = scene.philox_actions(np.arange(scene.n), sc0, k)
            parts.append(Trajectories(sink["j1"].numpy().copy(), sink["j2"].numpy().copy(), act[..., 0], act[..., 1], reward, flags))
        else:
            hact = torch.empty(k, scene.n, 2, dtype=torch.float32, pin_memory=True)
            hact.copy_(torch.as_tensor(np.asarray(actions[t:t + k], dtype=np.float32)))
            hout = scene.alloc_records(k, pinned_host=True)
            scene.rollout_host(k, hact, hout, chunk_steps=pipeline_steps)
            parts.append(Trajectories(hout["j1"].numpy().copy(), hout["j2"].numpy().copy(), hact[..., 0].numpy().copy(),
                                      hact[..., 1].numpy().copy(), hout["reward"].numpy().copy(), hout["flags"].numpy().copy()))
        t += k
    return Trajectories.concatenate(parts)


def run_experiment_exact(scene, steps, actions=None, generator=None, reset_u=None):
    """The same loop one step per launch (BatchedScene.step, then reset of the envs that terminated), keeping
    float64 joints and float64 actions: reproduces the reference's records digit for digit.
    actions: [steps,N,2] float64 or None (device uniforms)."""
    import torch
    K, N = int(steps), scene.n
    out = {f: np.zeros((K, N), dtype=np.float64) for f in ("j1", "j2", "a0", "a1", "reward")}
    flags = np.zeros((K, N), dtype=np.uint8)
    for t in range(K):
        if actions is None:
            act = scene.sample_action(0.1, generator=generator)
        else:
            act = torch.as_tensor(np.asarray(actions[t], dtype=np.float64), device=scene.device)
        j1, j2, rw, done, coll = scene.step(act)
        out["j1"][t], out["j2"][t] = j1.cpu().numpy(), j2.cpu().numpy()
        a = act.cpu().numpy()
        out["a0"][t], out["a1"][t] = a[:, 0], a[:, 1]
        out["reward"][t] = rw.cpu().numpy()
        flags[t] = scene.flags.cpu().numpy()
        term = scene.flags != 0
        if bool(term.any()):
            scene.reset(mask=term, reset_u=reset_u)            # experiment_0.py:30-34
    return Trajectories(out["j1"], out["j2"], out["a0"], out["a1"], out["reward"], flags)
