"""Lens factory (mirror of the reference's dff/factory.py:4-31; get_dataset is out of scope)."""
from deeplens.psfnet import PSFNet, ThinLens


def _make(args, split):
    cfg = args[split]
    ks, res, device = args['ks'], args['res'], args['device']
    if cfg['lens'] == 'thinlens':
        size = [float(i) for i in cfg['sensor_size']]
        return ThinLens(foc_len=cfg['foc_len'], fnum=cfg['fnum'], kernel_size=ks, sensor_size=size,
                        sensor_res=res).to(device)
    lens = PSFNet(filename=cfg['lens'], sensor_res=res, kernel_size=ks, device=device)
    lens.load_net(cfg['psfnet_path'])
    return lens


def get_lens(args):
    """args: the YAML dict of configs/aber_aware_dff_*.yml -> (train_lens, test_lens)."""
    return _make(args, 'train'), _make(args, 'test')
