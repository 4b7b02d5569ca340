"""pytest plug-in (`-p b200_pytest_plugin` with scripts/ on PYTHONPATH): installs the B200 interpolator subclasses in
the reference's interpolator_map before the reference's own tests are collected, so that its whole suite runs THROUGH
the plug-in classes (every test that does not ask for solver='b200' must behave exactly as with the stock classes)."""


def pytest_configure(config):
    import loopstructural_b200.plugin as b200

    classes = b200.install()
    config._b200_classes = classes


def pytest_report_header(config):
    return "loopstructural_b200 plug-in installed: " + ", ".join(c.__name__ for c in config._b200_classes.values())
