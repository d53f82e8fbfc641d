"""tools/vqgan/extract_vq.py — bulk audio -> VQ codes with the CUDA codec (same CLI as the reference tool).

    python tools/vqgan/extract_vq.py data/ --num-workers 8 --batch-size 64 \\
        --config-name modded_dac_vq --checkpoint-path checkpoints/s2-pro/codec.pth

`--num-workers N` re-spawns this script once per worker with CUDA_VISIBLE_DEVICES / SLURM_PROCID / SLURM_NTASKS set
(the reference's launcher, extract_vq.py:161-193); under SLURM or torchrun the existing rank variables are used."""
import os
import subprocess as sp
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))

import click

from fish_speech_b200.bulk_encode import logger, run_worker, worker_identity


def spawn_plan(num_workers: int, visible: list[str]) -> list[dict]:
    """Environment overrides of the N child workers: GPUs are dealt round-robin."""
    return [{"CUDA_VISIBLE_DEVICES": str(visible[i % len(visible)]), "SLURM_PROCID": str(i), "SLURM_NTASKS": str(num_workers)}
            for i in range(num_workers)]


@click.command()
@click.argument("folder")
@click.option("--num-workers", default=1)
@click.option("--config-name", default="modded_dac_vq")
@click.option("--checkpoint-path", default="checkpoints/s2-pro/codec.pth")
@click.option("--batch-size", default=64)
@click.option("--filelist", default=None, type=Path)
def main(folder: str, num_workers: int, config_name: str, checkpoint_path: str, batch_size: int, filelist: Path):
    _, world = worker_identity()
    if num_workers > 1 and world != num_workers:
        assert world == 1, "You should either use SLURM / torchrun or this launcher, not both"
        import torch

        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        visible = visible.split(",") if visible else [str(i) for i in range(max(1, torch.cuda.device_count()))]
        logger.info(f"Spawning {num_workers} workers")
        procs = []
        for over in spawn_plan(num_workers, visible):
            env = os.environ.copy()
            env.update(over)
            procs.append(sp.Popen([sys.executable] + sys.argv.copy(), env=env))
        rc = [p.wait() for p in procs]
        logger.info("All workers finished")
        if any(rc):
            raise SystemExit(max(rc))
        return
    run_worker(folder, config_name, checkpoint_path, batch_size, filelist)


if __name__ == "__main__":
    main()
