//! `Point`: what the facade needs from a 3-D vector type. Same required items as the reference trait
//! (constructor, three getters, three mutable getters) so existing impls for user types keep compiling; the
//! vector algebra the reference's CPU kernels used on it now lives on the device.
pub trait Point: Sized + Copy + Sync + Send + core::fmt::Debug + PartialEq {
    /// With the `serde` feature a point is serializable; implementors set `type Serde = Self;` (reference
    /// `src/point.rs:22-40`).
    #[cfg(feature = "serde")]
    type Serde: serde::Serialize + serde::de::DeserializeOwned;

    fn new(x: f32, y: f32, z: f32) -> Self;
    fn x(&self) -> f32;
    fn y(&self) -> f32;
    fn z(&self) -> f32;
    fn x_mut(&mut self) -> &mut f32;
    fn y_mut(&mut self) -> &mut f32;
    fn z_mut(&mut self) -> &mut f32;

    /// Component by index; panics like the reference for i > 2.
    fn get(&self, i: usize) -> f32 {
        [self.x(), self.y(), self.z()].get(i).copied().expect("Index out of bounds")
    }
    fn add(&self, o: &Self) -> Self { Self::new(self.x() + o.x(), self.y() + o.y(), self.z() + o.z()) }
    fn sub(&self, o: &Self) -> Self { Self::new(self.x() - o.x(), self.y() - o.y(), self.z() - o.z()) }
    fn dot(&self, o: &Self) -> f32 { self.x() * o.x() + self.y() * o.y() + self.z() * o.z() }
    fn cross(&self, o: &Self) -> Self {
        Self::new(self.y() * o.z() - self.z() * o.y(), self.z() * o.x() - self.x() * o.z(), self.x() * o.y() - self.y() * o.x())
    }
    fn length(&self) -> f32 { self.dot(self).sqrt() }
    fn dist(&self, o: &Self) -> f32 { self.sub(o).length() }
    fn dist2(&self, o: &Self) -> f32 { let d = self.sub(o); d.dot(&d) }
    fn fmul(&self, k: f32) -> Self { Self::new(self.x() * k, self.y() * k, self.z() * k) }
    fn comp_div(&self, o: &Self) -> Self { Self::new(self.x() / o.x(), self.y() / o.y(), self.z() / o.z()) }
}

macro_rules! impl_point_fields {
    ($t:ty, $ctor:expr) => {
        impl Point for $t {
            #[cfg(feature = "serde")]
            type Serde = Self;
            fn new(x: f32, y: f32, z: f32) -> Self { $ctor(x, y, z) }
            fn x(&self) -> f32 { self.x }
            fn y(&self) -> f32 { self.y }
            fn z(&self) -> f32 { self.z }
            fn x_mut(&mut self) -> &mut f32 { &mut self.x }
            fn y_mut(&mut self) -> &mut f32 { &mut self.y }
            fn z_mut(&mut self) -> &mut f32 { &mut self.z }
        }
    };
}

impl Point for [f32; 3] {
    #[cfg(feature = "serde")]
    type Serde = Self;
    fn new(x: f32, y: f32, z: f32) -> Self { [x, y, z] }
    fn x(&self) -> f32 { self[0] }
    fn y(&self) -> f32 { self[1] }
    fn z(&self) -> f32 { self[2] }
    fn x_mut(&mut self) -> &mut f32 { &mut self[0] }
    fn y_mut(&mut self) -> &mut f32 { &mut self[1] }
    fn z_mut(&mut self) -> &mut f32 { &mut self[2] }
}

#[cfg(feature = "glam")]
impl_point_fields!(glam::Vec3, glam::Vec3::new);
#[cfg(feature = "cgmath")]
impl_point_fields!(cgmath::Vector3<f32>, cgmath::Vector3::new);
#[cfg(feature = "mint")]
impl_point_fields!(mint::Vector3<f32>, |x, y, z| mint::Vector3 { x, y, z });
#[cfg(feature = "mint")]
impl_point_fields!(mint::Point3<f32>, |x, y, z| mint::Point3 { x, y, z });
#[cfg(feature = "nalgebra")]
impl_point_fields!(nalgebra::Vector3<f32>, nalgebra::Vector3::new);
#[cfg(feature = "nalgebra")]
impl_point_fields!(nalgebra::Point3<f32>, nalgebra::Point3::new);
