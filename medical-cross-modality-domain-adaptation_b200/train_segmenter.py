"""Entry point mirroring the reference's train_segmenter.py (:22-79): source-only (MR) segmenter training
with the same configuration literals.  `python train_segmenter.py [--batch-size N] ...`
Real TFRecord input is a "next" row; without it the synthetic source is used.
The reference's `os.system('tensorboard ...')` side effect (:69-70) is intentionally dropped."""
import argparse
import logging

from . import parallel
from . import runtime as rt
from . import source_segmenter as drn

train_fid = "./lists/mr_train_list"
val_fid = "./lists/mr_val_list"
output_path = "./tmp_exps/mr_baseline"
restore = True  # the reference restores even on a first run and falls through to "start from beginning" (:28,73)
num_cls = 5
batch_size = 10
training_iters = 10            # train_segmenter.py:33
epochs = 5000
checkpoint_space = 1500
optimizer = 'adam'
cost_kwargs = {"cross_flag": True, "miu_cross": 1.0, "dice_flag": True, "miu_dice": 1.0, "regularizer": 1e-4}
opt_kwargs = {"learning_rate": 1e-3}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch-size", type=int, default=batch_size)
    ap.add_argument("--training-iters", type=int, default=training_iters)
    ap.add_argument("--epochs", type=int, default=epochs)
    ap.add_argument("--keep-prob", type=float, default=0.75)
    ap.add_argument("--conv-backend", default=None, choices=["auto", "simt", "tc3", "tc1"])
    ap.add_argument("--output-path", default=output_path)
    a = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    parallel.init_from_env()
    if a.conv_backend:
        rt.set_conv_backend(a.conv_backend)
    net = drn.Full_DRN(channels=3, n_class=num_cls, batch_size=a.batch_size, cost_kwargs=dict(cost_kwargs))
    print("Network has been built ...")
    trainer = drn.Trainer(net, train_list=[], val_list=[], num_cls=num_cls, batch_size=a.batch_size, opt_kwargs=dict(opt_kwargs),
                          checkpoint_space=checkpoint_space, optimizer=optimizer, lr_update_flag=False)
    print("Now start training...")
    return trainer.train(output_path=a.output_path, restored_path=a.output_path, training_iters=a.training_iters,
                         epochs=a.epochs, restore=restore, dropout=a.keep_prob)


if __name__ == "__main__":
    main()
