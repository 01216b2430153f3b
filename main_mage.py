#!/usr/bin/env python
"""`python main_mage.py --split test --test_model <dir>/model_best.pth` -- the reference's sampling entry
(/root/reference/main_mage.py:29-56 flags, :201-248 sampling(), :276-297 dispatch) on the B200 path.

Kept from the reference: every flag; the yaml saved next to the checkpoint (`<dir>/config.yaml`) is what gets
loaded (main_mage.py:203); `instantiate_from_config(configs.model)`; checkpoint `{'state_dict': ...}` with an
optional DDP `module.` prefix (:218-223); a missing checkpoint file only prints a notice (:227-228);
`model.eval()`, `torch.no_grad()`, `autoregressive_generate(batch)` then `clamp_(-1, 1)` per sample (:240-242).

Additive flags (the reference hard-wires batch_size=1 and needs the datasets on disk):
  --batch-size B     prompts per generate call (per GPU)
  --synthetic N      N seeded synthetic prompts with the reference's batch-dict contract instead of `configs.data`
  --out DIR          write each clip as <video_id>.npy (the reference's GIF writer is commented out, :244-245)
  --config           used only when <dir>/config.yaml does not exist
Under torchrun (WORLD_SIZE > 1) the prompt list is cut into contiguous per-rank slices (mage_b200/shard.py);
there is no collective on the data path.  `--split train` is the stage-2 training driver, outside this path.

  --split val        (additive) the FORWARD half of that driver: the reference's periodic validation loss (main_mage.py:163-182) of
                     the checkpoint over the test split -- MAGE.forward in eval mode per batch, mean per rank, all_reduce(SUM) /
                     world size over NCCL.  Needs clips with all frames_length frames and a checkpoint that holds the video
                     posterior (the model is built with `with_posterior`).  --seed fixes the reparameterisation draws.
"""
import argparse
import os
import sys
import time
from collections import OrderedDict

import torch
from torch.utils.data import DataLoader, Subset

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mage_b200 import shard  # noqa: E402
from mage_b200.config import load_yaml  # noqa: E402
from utils.util import instantiate_from_config  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--config', type=str, default='config/mage_caterv2_L32.yaml')
parser.add_argument('--split', type=str, default='test')
parser.add_argument('--checkpoint-path', type=str, default='../results//')
parser.add_argument('--device', type=str, default='cuda')
parser.add_argument("--num-workers", type=int, default=4)
parser.add_argument('--world-size', default=-1, type=int, help='number of nodes for distributed training')
parser.add_argument('--rank', default=-1, type=int, help='node rank for distributed training')
parser.add_argument('--dist-url', default='tcp://127.0.0.1:65532', type=str, help='url used to set up distributed training')
parser.add_argument('--dist-backend', default='nccl', type=str, help='distributed backend')
parser.add_argument('--seed', default=None, type=int, help='seed for initializing training. ')
parser.add_argument('--gpu', default=0, type=int, help='GPU id to use.')
parser.add_argument('--multiprocessing-distributed', action='store_true')
parser.add_argument("--n_samples", type=int, default=1, help="how many samples to produce for each instance")
parser.add_argument("--test_model", type=str, default='./models/MAGE/catergenv2/model_best.pth')
# additive
parser.add_argument("--batch-size", type=int, default=1, help="prompts per generate call (reference: 1)")
parser.add_argument("--synthetic", type=int, default=0, help="use N seeded synthetic prompts instead of configs.data")
parser.add_argument("--out", type=str, default=None, help="directory for the generated clips (.npy)")
parser.add_argument("--caption", type=str, default=None, help="sample ONE clip from this caption (needs --image); word-level vocabulary of --dataset")
parser.add_argument("--image", type=str, default=None, help="first frame for --caption (any PIL-readable file)")
parser.add_argument("--speed", type=float, default=0.5, help="speed in [0,1) for --caption")
parser.add_argument("--dataset", type=str, default=None, choices=["mnist", "caterv1", "caterv2"],
                    help="vocabulary for --caption (default: configs.data.params.dataset, else by vocab size)")
parser.add_argument("--gifs", action="store_true", help="also write <ckpt dir>/videos/<video_id>.gif like the reference's save_gifs")


def load_configs(opt) -> dict:
    beside = os.path.join(os.path.dirname(opt.test_model), "config.yaml")
    return load_yaml(beside if os.path.isfile(beside) else opt.config)


def load_checkpoint(model, opt) -> bool:
    """main_mage.py:210-228."""
    test_model = opt.test_model
    if not os.path.isfile(test_model):
        print("=> no checkpoint found at '{}'".format(test_model))
        return False
    loc = None if opt.gpu is None else 'cuda:{}'.format(opt.gpu) if torch.cuda.is_available() else 'cpu'
    checkpoint = torch.load(test_model, map_location=loc)
    sd = checkpoint['state_dict']
    if list(sd.keys())[0].startswith('module.'):
        sd = OrderedDict((k[7:], v) for k, v in sd.items())
    model.load_state_dict(sd)
    print("=> loaded checkpoint '{}'".format(test_model))
    return True


def save_gifs(tgr, video_id, test_model):
    """main_mage.py:250-257 (commented out at its call site :244-245): frames [L,C,H,W] in [-1,1] -> <ckpt dir>/videos/<id>.gif at
    3 fps.  The reference uses imageio; PIL writes the same frames here."""
    import numpy as np
    from PIL import Image
    tgr_imgs = (tgr + 1) * 0.5
    tgr_imgs = (tgr_imgs * 255.).numpy().astype(np.uint8).transpose(0, 2, 3, 1)
    save_path = os.path.join(os.path.dirname(test_model), 'videos')
    if not os.path.exists(save_path):
        os.makedirs(save_path)
    frames = [Image.fromarray(f[..., 0] if f.shape[-1] == 1 else f) for f in tgr_imgs]
    path = os.path.join(save_path, video_id + '.gif')
    frames[0].save(path, save_all=True, append_images=frames[1:], duration=1000 // 3, loop=0)
    return path


def sampling(opt):
    configs = load_configs(opt)
    rank, world, local_rank = shard.env_world()
    if world > 1:
        opt.gpu = local_rank
    if torch.cuda.is_available():
        torch.cuda.set_device(opt.gpu or 0)
    device = torch.device(opt.device if opt.device != 'cuda' else 'cuda:{}'.format(opt.gpu or 0))
    if world > 1:
        shard.init_distributed(opt.dist_backend, device if device.type == 'cuda' else None)

    if opt.caption is not None:
        from dataload import VOCABS, encode_caption, load_first_frame
        mp = configs['model']['params']
        fs = mp['first_stage_config']['params']
        ds = opt.dataset or (configs.get('data', {}).get('params', {}) or {}).get('dataset')
        if ds not in VOCABS:
            sizes = {len(v): k for k, v in VOCABS.items() if k != 'caterv1'}
            ds = sizes.get(mp['text_encoder_config']['params']['vocab_size'], 'mnist' if fs['down_ratio'] == 4 else 'caterv2')
        assert opt.image, "--caption needs --image (the clip's first frame)"
        item = {'images': load_first_frame(opt.image, fs['input_dim'], mp['image_resolution'] * fs['down_ratio']),
                'text': encode_caption(opt.caption, ds), 'speed': torch.tensor(opt.speed, dtype=torch.float32), 'video_id': 'caption_0'}
        test_dataset = [item]
    elif opt.synthetic > 0:
        from dataload import SyntheticCaptionVideos
        test_dataset = SyntheticCaptionVideos(configs['model']['params'], opt.synthetic, seed=1234 if opt.seed is None else opt.seed)
    else:
        test_dataset = instantiate_from_config(configs['data'], {'split': 'test'})
    lo, hi = shard.shard_bounds(len(test_dataset), world, rank)
    from dataload import collate_fn
    test_dataloader = DataLoader(Subset(test_dataset, range(lo, hi)), batch_size=opt.batch_size, shuffle=False, num_workers=0,
                                 pin_memory=torch.cuda.is_available(), collate_fn=collate_fn)

    model = instantiate_from_config(configs['model'])
    model = model.to(device)
    load_checkpoint(model, opt)
    model.eval()
    if opt.out:
        os.makedirs(opt.out, exist_ok=True)
    frames = 0
    t0 = time.perf_counter()
    with torch.no_grad():
        idx = 0
        seen = 0
        for batch in test_dataloader:
            video_ids = batch.pop('video_id', None)
            for k in batch.keys():
                batch[k] = batch[k].to(device)
            n_here = batch['text'].shape[0]
            for rep in range(opt.n_samples):
                noise = None
                if opt.seed is not None and getattr(model, 'randomness', False):
                    # additive: with --seed a prompt's AdaIN noise depends only on (seed, sample repetition, GLOBAL prompt index), so
                    # the clips do not depend on the batch size or on how many GPUs share the prompt list; without --seed the noise
                    # comes from the process's default CPU generator inside the call, exactly like the reference (mage_model.py:661)
                    noise = shard.noise_for_prompts(opt.seed + 7919 * rep, range(lo + seen, lo + seen + n_here), model.image_resolution)
                generated = model.autoregressive_generate(batch) if noise is None else model.autoregressive_generate(batch, noise=noise)
                generated.clamp_(min=-1, max=1)
                frames += generated.shape[0] * (generated.shape[1] - 1)
            if opt.out or opt.gifs:
                import numpy as np
                clips = generated.cpu()
                for b in range(clips.shape[0]):
                    name = video_ids[b] if video_ids else f"r{rank}_{idx}_{b}"
                    if opt.out:
                        np.save(os.path.join(opt.out, name + ".npy"), clips[b].numpy())
                    if opt.gifs:
                        save_gifs(clips[b], name, opt.test_model)
            print(idx)
            idx += 1
            seen += n_here
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    if world > 1:
        secs = shard.max_over_ranks(secs, device)
        frames = shard.sum_over_ranks(frames, device)
    if rank == 0:
        print(f"generated {int(frames)} frames in {secs:.2f}s on {world} GPU(s)")
    return frames


def validation(opt):
    """main_mage.py:163-182 for one checkpoint: `test_loss` of the test split, printed by rank 0."""
    configs = load_configs(opt)
    rank, world, local_rank = shard.env_world()
    if world > 1:
        opt.gpu = local_rank
    torch.cuda.set_device(opt.gpu or 0)
    device = torch.device('cuda:{}'.format(opt.gpu or 0))
    if world > 1:
        shard.init_distributed(opt.dist_backend, device)
    mp = configs['model']['params']
    L = mp['frames_length']
    if opt.synthetic > 0:
        from dataload import SyntheticCaptionVideos
        test_dataset = SyntheticCaptionVideos(mp, opt.synthetic, seed=1234 if opt.seed is None else opt.seed, frames=L)
    else:
        test_dataset = instantiate_from_config(configs['data'], {'split': 'test'})
    lo, hi = shard.shard_bounds(len(test_dataset), world, rank)
    from dataload import collate_fn
    loader = DataLoader(Subset(test_dataset, range(lo, hi)), batch_size=opt.batch_size, shuffle=False, num_workers=0,
                        pin_memory=True, collate_fn=collate_fn)
    cfg = dict(configs['model'])
    cfg['params'] = dict(mp, with_posterior=True)
    model = instantiate_from_config(cfg).to(device)
    load_checkpoint(model, opt)
    model.eval()
    seen = [0]

    def loss_fn(batch):
        eps = None
        if opt.seed is not None and model.randomness:   # the draw of a clip depends only on (seed, global clip index)
            n = batch['text'].shape[0]
            eps = shard.noise_for_prompts(opt.seed, range(lo + seen[0], lo + seen[0] + n), model.image_resolution)
            seen[0] += n
        return model(batch, eps=eps)

    t0 = time.perf_counter()
    test_loss = shard.validation_loss(loss_fn, loader, device)
    torch.cuda.synchronize()
    if rank == 0:
        print("test_loss = %.6f  (%d clips of %d frames, %d GPU(s), %.2fs)" % (test_loss, len(test_dataset), L, world, time.perf_counter() - t0))
    return test_loss


if __name__ == "__main__":
    opt = parser.parse_args()
    if opt.split == 'test':
        sampling(opt)
    elif opt.split == 'val':
        validation(opt)
    elif opt.split == 'train':
        raise SystemExit("--split train (stage-2 training, main_mage.py:58-199) is outside the sampling path this repo implements "
                         "(SURVEY.md §8f N2)")
    else:
        raise SystemExit(f"unknown --split {opt.split!r}")
